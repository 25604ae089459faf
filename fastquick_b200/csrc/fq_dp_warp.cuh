// Warp-cooperative dynamic programming for rows a10/a11 (device only): one alignment per WARP.
// Two forms.  The WAVEFRONT form (fq_dp_wave.cuh: a lane owns consecutive columns in registers, rows skewed over the lanes)
// carries the forward pass of the local alignment and the banded global alignments whenever the window fits
// (<= 768 columns / <= 255 columns).  The ROW-CHUNK form below (the 32 lanes sweep a row in chunks of 32 columns; the gap that
// extends along the row -- F of the local pass, D of the global pass -- is a (max,+) prefix scan over the lanes, exact in
// integer arithmetic; rows in shared memory, trace-back in a warp-private slab of global memory) carries the reverse pass
// of the local alignment, whose band limits depend on the running maximum of the rows before, and everything wider.
// Same recurrences, boundary rules and tie-breaks as fq_device_dp.cuh (which remains the
// host-checkable statement of the logic and the fallback for windows that do not fit).
#pragma once
#include "fq_device_dp.cuh"
#include "fq_dp_wave.cuh"

namespace fqb {

#define FQB_FULL 0xffffffffu

struct WarpDp {
    int32_t *sm; int n_ints;          // per-warp shared-memory rows
    uint8_t *refc; int n_refc;        // per-warp shared-memory copy of the reference window, one nt4 code per byte
    uint8_t *ops; int n_ops;          // per-warp shared memory: the path ops of the last global alignment
    uint8_t *gb; int n_bytes;         // per-warp global slab: trace matrix of the row-chunk global alignment
    int lane;
};

__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FQB_FULL, v, d); if (lane >= d) v = t > v ? t : v; }
    return v;
}

constexpr int kVeryNeg = -2000000000;

// unpack the 2-bit reference window once per alignment; the DP loops then read one shared-memory byte per cell
__device__ __forceinline__ bool warp_load_ref(const RefWin &R, const WarpDp &w) {
    if (R.l > w.n_refc) return false;
    for (int i = w.lane; i < R.l; i += 32) w.refc[i] = (uint8_t)R.at(i);
    __syncwarp();
    return true;
}
// aln_sm_maq with a reference base that is never N: row-constant part hoisted (qn = read base is N)
__device__ __forceinline__ int maq_row_score(uint32_t a, uint32_t qj, bool qn) { return qn ? -13 : (a == qj ? 11 : -19); }

// aln_global_core, row-parallel.  Result broadcast to all lanes; path ops in w.ops[0 .. n_ops).
__device__ GlobalResult warp_global_align(const RefWin &R, int r0, int len1, const ReadSeq &Q, int q0, int len2, int gap_end, int band,
                                          const WarpDp &w) {
    GlobalResult res; res.score = 0; res.n_ops = 0; res.too_big = false;
    if (len1 == 0 || len2 == 0) return res;
    const int lane = w.lane;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; } else { b1 = band; b2 = len2 - len1 + band; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const int W = len1 + 1;
    const int tw = (b1 + b2 <= len1) ? (b1 + b2 + 1) : W;
    const int end_ge = gap_end >= 0 ? gap_end : kGapExt;
    const int ops_cap = len1 + len2 + 2;
    if (6 * W > w.n_ints || (len2 + 1) * tw > w.n_bytes || ops_cap > w.n_ops) { res.too_big = true; return res; }
    int32_t *sm = w.sm;
    uint8_t *tr = w.gb;
#define WROW(rw, which, i) sm[((rw) * 3 + (which)) * W + (i)]
#define WTRC(j, i) tr[(j) * tw + ((i) - ((j) > b2 ? (j) - b2 : 0))]
    int cur = 0, last = 1;
    for (int i = lane; i < b1; i += 32) {                // first row: D chain with the end-gap extension
        WROW(cur, 0, i) = i == 0 ? 0 : kNegInf;
        WROW(cur, 1, i) = kNegInf;
        WROW(cur, 2, i) = i == 0 ? kNegInf : -(kGapOpen + i * end_ge);
        if (i) WTRC(0, i) = (uint8_t)((i == 1 ? kOpM : kOpD) << 4);
    }
    __syncwarp();
    { int t = cur; cur = last; last = t; }
    const int tmp_end = (b2 < len2) ? b2 : len2 - 1;
    for (int j = 1; j <= len2; ++j) {
        const bool head = j <= tmp_end || (j == tmp_end + 1 && j == len2 && b2 != len2 - 1);
        const bool mid = !head && j <= len2 - b2 + 1;
        const bool last_row_d = head ? (j == tmp_end + 1) : (!mid && j == len2);
        const int d_ge = last_row_d ? end_ge : kGapExt;
        const uint32_t qj = Q.at(q0 + j - 1);
        const bool qn = qj > 3;
        int first, endc;
        int bM = kNegInf, bD = kNegInf;                  // boundary cell of this row (column `first`)
        if (head) {
            first = 0;
            endc = (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1;
            if (lane == 0) {
                int iv; uint32_t t;
                dp_gap(WROW(last, 0, 0), WROW(last, 1, 0), kGapOpen, end_ge, kOpI, iv, t);
                WROW(cur, 0, 0) = kNegInf; WROW(cur, 1, 0) = iv; WROW(cur, 2, 0) = kNegInf;
                WTRC(j, 0) = (uint8_t)(t << 2);
            }
        } else {
            first = j - b2;
            endc = mid ? j + b1 - 1 : len1;
            if (lane == 0) { WROW(cur, 0, first) = kNegInf; WROW(cur, 1, first) = kNegInf; WROW(cur, 2, first) = kNegInf; }
        }
        // the last column of a row: whether the cell above exists, and set_end_I's extension penalty (row constants)
        const bool last_have_up = head ? (j + b1 - 1 > len1) : !mid;
        const int last_ige = (head || !mid) ? end_ge : kGapExt;
        int carryM = bM, carryD = bD, carryP = bD + first * d_ge;
        for (int base = first + 1; base <= endc; base += 32) {
            const int i = base + lane;
            const bool on = i <= endc;
            int m = kVeryNeg, iv = kNegInf, d;
            uint32_t tm = 0, ti = 0, td;
            if (on) {
                const int sco = maq_row_score(w.refc[r0 + i - 1], qj, qn);
                dp_from_diag(WROW(last, 0, i - 1), WROW(last, 1, i - 1), WROW(last, 2, i - 1), sco, m, tm);
                const bool lastc = i == endc;
                if (!lastc || last_have_up)
                    dp_gap(WROW(last, 0, i), WROW(last, 1, i), kGapOpen, lastc ? last_ige : kGapExt, kOpI, iv, ti);
            }
            // D(i) = max(M(i-1) - go, D(i-1)) - ge  ==  max_k<i (M(k) - go + k ge) - i ge   (prefix max over the row)
            int mprev = __shfl_up_sync(FQB_FULL, m, 1);
            if (lane == 0) mprev = carryM;
            int b = on ? mprev - kGapOpen + (i - 1) * d_ge : kVeryNeg;
            int pm = warp_incl_max(b, lane);
            if (carryP > pm) pm = carryP;
            d = pm - i * d_ge;
            int dprev = __shfl_up_sync(FQB_FULL, d, 1);
            if (lane == 0) dprev = carryD;
            td = (mprev - kGapOpen > dprev) ? kOpM : kOpD;
            if (on) {
                WROW(cur, 0, i) = m; WROW(cur, 1, i) = iv; WROW(cur, 2, i) = d;
                WTRC(j, i) = (uint8_t)(tm | ti << 2 | td << 4);
            }
            const int lastl = (endc - base) < 31 ? (endc - base) : 31;
            carryM = __shfl_sync(FQB_FULL, m, lastl);
            carryD = __shfl_sync(FQB_FULL, d, lastl);
            carryP = __shfl_sync(FQB_FULL, pm, lastl);
        }
        __syncwarp();
        { int t = cur; cur = last; last = t; }
    }
    int score = 0, n_ops = 0;
    if (lane == 0) {                                     // back-trace (stdaln.c:480-512)
        int i = len1, j = len2;
        int mx = WROW(last, 0, len1);
        uint32_t cell = WTRC(j, i), type = cell & 3, ctype = kOpM;
        if (WROW(last, 1, len1) > mx) { mx = WROW(last, 1, len1); type = (cell >> 2) & 3; ctype = kOpI; }
        if (WROW(last, 2, len1) > mx) { mx = WROW(last, 2, len1); type = (cell >> 4) & 3; ctype = kOpD; }
        int n = 0;
        w.ops[n++] = (uint8_t)ctype;
        do {
            if (ctype == kOpM) { --i; --j; } else if (ctype == kOpI) --j; else --i;
            cell = (i == 0 && j == 0) ? 0 : WTRC(j, i);
            ctype = type;
            type = ctype == kOpM ? (cell & 3) : ctype == kOpI ? ((cell >> 2) & 3) : ((cell >> 4) & 3);
            w.ops[n++] = (uint8_t)ctype;
        } while (i || j);
        score = mx; n_ops = n - 1;
    }
    res.score = __shfl_sync(FQB_FULL, score, 0);
    res.n_ops = __shfl_sync(FQB_FULL, n_ops, 0);
    __syncwarp();
#undef WROW
#undef WTRC
    return res;
}

// the banded global alignment of a window whose codes sit in w.refc: wavefront form when the columns fit the registers
// of the lanes and the trace-back fits the shared-memory rows (which are free at that point), row-chunk form otherwise
__device__ GlobalResult warp_global_any(const RefWin &R, int r0, int len1, const ReadSeq &Q, int q0, int len2, int gap_end, int band, const WarpDp &w) {
    // shared-memory rows: len2 + 1 words of per-row constants, then the trace-back (64 or 128 bytes per row)
    const int avail = w.n_ints * 4 - (len2 + 1) * 4;
    if (len1 <= 32 * 4 && len2 * 64 <= avail)
        return wave_global_align<4, uint16_t>(w.refc, r0, len1, Q, q0, len2, gap_end, band, w.sm, reinterpret_cast<uint16_t *>(w.sm + len2 + 1), len2, w.ops, w.n_ops, w.lane);
    if (len1 <= 32 * 5 && len2 * 128 <= avail)       // the fewest columns per lane that fit 32 lanes: a step costs 30 + 37 CT instructions
        return wave_global_align<5, uint32_t>(w.refc, r0, len1, Q, q0, len2, gap_end, band, w.sm, reinterpret_cast<uint32_t *>(w.sm + len2 + 1), len2, w.ops, w.n_ops, w.lane);
    if (len1 <= 32 * 6 && len2 * 128 <= avail)
        return wave_global_align<6, uint32_t>(w.refc, r0, len1, Q, q0, len2, gap_end, band, w.sm, reinterpret_cast<uint32_t *>(w.sm + len2 + 1), len2, w.ops, w.n_ops, w.lane);
    if (len1 <= 32 * 8 && len2 * 128 <= avail)
        return wave_global_align<8, uint32_t>(w.refc, r0, len1, Q, q0, len2, gap_end, band, w.sm, reinterpret_cast<uint32_t *>(w.sm + len2 + 1), len2, w.ops, w.n_ops, w.lane);
    return warp_global_align(R, r0, len1, Q, q0, len2, gap_end, band, w);
}

// aln_local_core (_thres = 1): forward and reverse passes row-parallel with an F prefix scan, then the global
// alignment of the local region.  Previous-row state in shared memory: H[i] = h(i, previous row), E[i] = e(i, previous row).
// All lanes return the same result; the path ops are left in w.ops[0 .. n_ops).
__device__ LocalResult warp_local_align(const RefWin &R, int len1, const ReadSeq &Q, int len2, const WarpDp &w) {
    LocalResult res; res.score = -1; res.n_ops = 0; res.start_i = res.start_j = res.end_i = res.end_j = 0; res.too_big = false;
    if (len1 == 0 || len2 == 0) return res;
    const int lane = w.lane, q = kGapOpen, r = kGapExt, qr = q + r, max_score = 11, W = len1 + 2;
    if (2 * W > w.n_ints || !warp_load_ref(R, w)) { res.too_big = true; return res; }
    int32_t *H = w.sm, *E = w.sm + W;
    for (int i = lane; i < W; i += 32) { H[i] = 0; E[i] = 0; }
    __syncwarp();
    int score_f = 0, end_i = 0, end_j = 0;
    if (len1 <= 32 * 8) wave_local_forward<8>(w.refc, len1, Q, len2, lane, score_f, end_i, end_j);
    else if (len1 <= 32 * 16) wave_local_forward<16>(w.refc, len1, Q, len2, lane, score_f, end_i, end_j);
    else if (len1 <= 32 * 24) wave_local_forward<24>(w.refc, len1, Q, len2, lane, score_f, end_i, end_j);
    else
    for (int j = 1; j <= len2; ++j) {
        const uint32_t qj = Q.at(j - 1);
        const bool qn = qj > 3;
        int carry_g = kVeryNeg;               // max over the columns of earlier chunks of h'(k) + k r
        int carry_diag = 0;                   // h(base - 1, j - 1); column 0 holds 0
        int row_best = 0, row_best_i = 0;
        for (int base = 1; base <= len1; base += 32) {
            const int i = base + lane;
            const bool on = i <= len1;
            const int hp = on ? H[i] : 0, ep = on ? E[i] : 0;
            int hd = __shfl_up_sync(FQB_FULL, hp, 1);
            if (lane == 0) hd = carry_diag;
            int e = 0;
            if (hp >= qr + 1) { e = ep - r; if (hp - qr > e) e = hp - qr; }
            int h1 = on ? hd + maq_row_score(w.refc[i - 1], qj, qn) : 0;
            if (h1 < 0) h1 = 0;
            if (h1 < e) h1 = e;
            // f(i) = max_{k < i} (h'(k) - q - (i - k) r): the gap along the row as a prefix maximum
            const int g = on ? h1 + i * r : kVeryNeg;
            const int pm = warp_incl_max(g, lane);
            int pe = __shfl_up_sync(FQB_FULL, pm, 1);
            if (lane == 0) pe = kVeryNeg;
            if (carry_g > pe) pe = carry_g;
            int h = h1;
            if (pe > kVeryNeg / 2) { const int f = pe - q - i * r; if (f > h) h = f; }
            const int lastl = (len1 - base) < 31 ? (len1 - base) : 31;
            carry_diag = __shfl_sync(FQB_FULL, hp, lastl);
            { const int t = __shfl_sync(FQB_FULL, pm, lastl); if (t > carry_g) carry_g = t; }
            __syncwarp();
            if (on) {
                H[i] = h; E[i] = e;
                if (h > row_best) { row_best = h; row_best_i = i; }     // ascending i per lane, strict > keeps the first
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {      // row maximum, smallest column attaining it
            const int ob = __shfl_xor_sync(FQB_FULL, row_best, d), oi = __shfl_xor_sync(FQB_FULL, row_best_i, d);
            if (ob > row_best || (ob == row_best && oi < row_best_i)) { row_best = ob; row_best_i = oi; }
        }
        if (row_best > score_f) { score_f = row_best; end_i = row_best_i; end_j = j; }
        __syncwarp();
    }
    res.score = score_f;
    if (score_f < 1) return res;
    for (int i = lane; i <= end_i; i += 32) { H[i] = 0; E[i] = 0; }
    __syncwarp();
    if (end_i == 0 || end_j == 0) return res;
    int score_r = maq_score(w.refc[end_i - 1], Q.at(end_j - 1));
    int start_i = end_i, start_j = end_j;
    if (lane == 0) H[end_i] = qr + score_r;
    __syncwarp();
    // reverse pass over the band (end, start], columns descending: lane L of a chunk owns column base - L.
    // A cell reads H[i+1] (diagonal), H[i] (row j+1) and E[i]; like the reference it then stores h of the cell
    // BEFORE it into H[i+1] (0 for the first cell of the row) and its own e into E[i].
    int start = end_i - 1, end = end_i - 3;
    if (end <= 0) end = 0;
    for (int j = end_j - 1; j != 0; --j) {
        const uint32_t qj = Q.at(j - 1);
        const bool qn = qj > 3;
        int carry_g = kVeryNeg, carry_h = 0, carry_hmax = kVeryNeg;
        int row_best = score_r, row_best_i = 0;
        int hit_i = 0;
        bool hit = false;
        for (int base = start; base > end; base -= 32) {
            const int i = base - lane;
            const bool on = i > end;
            const int hd = on ? H[i + 1] : 0, hu = on ? H[i] : 0, eo = on ? E[i] : 0;
            int h1 = on ? hd + maq_row_score(w.refc[i - 1], qj, qn) : 0;
            if (h1 < 0) h1 = 0;
            int e = eo - r; if (hu - qr > e) e = hu - qr;
            if (e < 0) e = 0;
            if (h1 < e) h1 = e;
            const int g = on ? h1 - i * r : kVeryNeg;           // f(i) = max_{k > i} (h'(k) - q - (k - i) r)
            const int pm = warp_incl_max(g, lane);
            int pe = __shfl_up_sync(FQB_FULL, pm, 1);
            if (lane == 0) pe = kVeryNeg;
            if (carry_g > pe) pe = carry_g;
            int h = h1;
            if (pe > kVeryNeg / 2) { const int f = pe - q + i * r; if (f > h) h = f; }
            int hnext = __shfl_up_sync(FQB_FULL, h, 1);
            if (lane == 0) hnext = carry_h;
            const int nact = (base - end) < 32 ? (base - end) : 32;
            carry_h = __shfl_sync(FQB_FULL, h, nact - 1);
            { const int t = __shfl_sync(FQB_FULL, pm, nact - 1); if (t > carry_g) carry_g = t; }
            __syncwarp();
            if (on) {
                H[i + 1] = hnext; E[i] = e;
                if (h > row_best) { row_best = h; row_best_i = i; }     // descending i per lane, strict > keeps the first
            }
            // the reference stops at a cell only if it is a NEW running maximum (cells are visited in descending i) that
            // equals score_f + qr; the reverse pass can exceed that value elsewhere in the row (the forward pass cuts
            // e-chains below q + r, the reverse pass does not), so the running maximum has to be reproduced exactly
            const int hx = warp_incl_max(on ? h : kVeryNeg, lane);
            int run = __shfl_up_sync(FQB_FULL, hx, 1);
            if (lane == 0) run = kVeryNeg;
            if (carry_hmax > run) run = carry_hmax;
            if (score_r > run) run = score_r;
            { const int t = __shfl_sync(FQB_FULL, hx, nact - 1); if (t > carry_hmax) carry_hmax = t; }
            const unsigned hm = __ballot_sync(FQB_FULL, on && h > run && h - qr == score_f);
            if (hm) { hit = true; hit_i = base - (__ffs(hm) - 1); break; }
        }
        if (hit) { score_r = score_f + qr; start_i = hit_i; start_j = j; break; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {      // row maximum, largest column attaining it
            const int ob = __shfl_xor_sync(FQB_FULL, row_best, d), oi = __shfl_xor_sync(FQB_FULL, row_best_i, d);
            if (ob > row_best || (ob == row_best && oi > row_best_i)) { row_best = ob; row_best_i = oi; }
        }
        if (row_best > score_r) { score_r = row_best; start_i = row_best_i; start_j = j; }
        if (lane == 0) { H[end + 1] = carry_h; E[end] = 0; }
        __syncwarp();
        if (H[start] <= qr) --start;
        if (start <= 0) start = 0;
        end = start_i - (start_j - j) - (score_r + (start_j - j) * max_score) / r - 1;
        if (end <= 0) end = 0;
        __syncwarp();
    }
    score_r -= qr;
    int jmax = (end_i - start_i > end_j - start_j) ? end_i - start_i : end_j - start_j;
    ++jmax;
    GlobalResult g;
    for (int bw = kBandWidth;; bw <<= 1) {
        g = warp_global_any(R, start_i - 1, end_i - start_i + 1, Q, start_j - 1, end_j - start_j + 1, -1, bw, w);
        if (g.too_big) { res.too_big = true; return res; }
        if (g.score == score_r || score_f == g.score) break;
        if (bw > jmax) break;
    }
    res.score = (score_r > g.score && score_f > g.score) ? -1 : g.score;
    res.n_ops = g.n_ops;
    res.start_i = start_i; res.start_j = start_j; res.end_i = end_i; res.end_j = end_j;
    return res;
}

// ---- very wide mate-rescue windows -----------------------------------------------------------------------------------
// (they arise when a gapped forward read's provisional position underflowed and bwa_paired_sw's window arithmetic wraps:
// the reference then runs aln_local_core over almost the whole reduced reference.)  The forward pass is split into column
// slices that are scanned independently: a local alignment of a len2-base read with a positive score spans fewer than
// len2 + 11 len2 / 9 columns, so a slice that starts kScanOverlap columns before its core computes exact H values on the
// core.  Each slice reports its first maximum in row-major order; the smallest (row, column) among the slices that reach the
// global maximum is the cell the reference's single pass would report.
constexpr int kScanCore = 1792, kScanOverlap = 256;       // columns per slice / overlap (>= 100 + 1100 / 9 for reads <= 100 bp... checked at run time)
struct ScanBest { int score, j, i; };

__device__ ScanBest warp_local_scan_slice(const RefWin &R, int core_lo, int core_hi, const ReadSeq &Q, int len2, const WarpDp &w) {
    ScanBest best; best.score = 0; best.j = 0; best.i = 0;
    const int lane = w.lane, q = kGapOpen, r = kGapExt, qr = q + r;
    const int c0 = core_lo - kScanOverlap > 1 ? core_lo - kScanOverlap : 1;     // first column computed (1-based in the window)
    const int wl = core_hi - c0 + 1, W = wl + 2;
    if (2 * W > w.n_ints || wl > w.n_refc) { best.score = -2; return best; }
    int32_t *H = w.sm, *E = w.sm + W;
    for (int i = lane; i < W; i += 32) { H[i] = 0; E[i] = 0; }
    for (int i = lane; i < wl; i += 32) w.refc[i] = (uint8_t)R.at(c0 - 1 + i);
    __syncwarp();
    const int first_core = core_lo - c0 + 1;                                   // local index of the first core column
    for (int j = 1; j <= len2; ++j) {
        const uint32_t qj = Q.at(j - 1);
        const bool qn = qj > 3;
        int carry_g = kVeryNeg, carry_diag = 0, row_best = 0, row_best_i = 0;
        for (int base = 1; base <= wl; base += 32) {
            const int i = base + lane;
            const bool on = i <= wl;
            const int hp = on ? H[i] : 0, ep = on ? E[i] : 0;
            int hd = __shfl_up_sync(FQB_FULL, hp, 1);
            if (lane == 0) hd = carry_diag;
            int e = 0;
            if (hp >= qr + 1) { e = ep - r; if (hp - qr > e) e = hp - qr; }
            int h1 = on ? hd + maq_row_score(w.refc[i - 1], qj, qn) : 0;
            if (h1 < 0) h1 = 0;
            if (h1 < e) h1 = e;
            const int g = on ? h1 + i * r : kVeryNeg;
            const int pm = warp_incl_max(g, lane);
            int pe = __shfl_up_sync(FQB_FULL, pm, 1);
            if (lane == 0) pe = kVeryNeg;
            if (carry_g > pe) pe = carry_g;
            int h = h1;
            if (pe > kVeryNeg / 2) { const int f = pe - q - i * r; if (f > h) h = f; }
            const int lastl = (wl - base) < 31 ? (wl - base) : 31;
            carry_diag = __shfl_sync(FQB_FULL, hp, lastl);
            { const int t = __shfl_sync(FQB_FULL, pm, lastl); if (t > carry_g) carry_g = t; }
            __syncwarp();
            if (on) {
                H[i] = h; E[i] = e;
                if (i >= first_core && h > row_best) { row_best = h; row_best_i = i; }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ob = __shfl_xor_sync(FQB_FULL, row_best, d), oi = __shfl_xor_sync(FQB_FULL, row_best_i, d);
            if (ob > row_best || (ob == row_best && oi < row_best_i)) { row_best = ob; row_best_i = oi; }
        }
        if (row_best > best.score) { best.score = row_best; best.j = j; best.i = c0 - 1 + row_best_i; }
        __syncwarp();
    }
    return best;
}

struct HugeJob {                      // one very wide window of one mate
    uint32_t pair; int32_t k;
    long long beg; int32_t reglen, strand;
    int32_t n_slices, first_slice;
};
struct HugeView { const HugeJob *jobs; const ScanBest *slice_best; int n_jobs; };

// bwa_sw_core by a warp: all lanes run the checks and the DP, lane 0 turns the path into the CIGAR and counts, and
// the values the caller's control flow depends on are broadcast (the other lanes' cigar[] holds only the end elements).
struct WarpSwCore {
    const WarpDp &w;
    const HugeView *huge;             // scan results for the windows that do not fit the shared-memory rows (or nullptr)
    uint32_t pair; int k_hint;        // which pair is being processed (to find its scan results); k_hint is advanced per call
    __device__ int operator()(int64_t l_pac, const uint8_t *pac, const ReadSeq &Q, int64_t *beg, int reglen, uint16_t *cigar, uint32_t *cnt) const {
        const int len = Q.len, lane = w.lane;
        if (reglen < 20 || l_pac - *beg < len) return 0;
        int nn = 0;
        for (int k = lane; k < len; k += 32) nn += Q.at(k) >= 4;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nn += __shfl_xor_sync(FQB_FULL, nn, d);
        if ((float)nn / len >= 0.25f || len - nn < 20) return 0;
        RefWin R; R.pac = pac; R.beg = *beg;
        { int64_t e = *beg + reglen < l_pac ? *beg + reglen : l_pac; R.l = (int)(e - *beg); }
        LocalResult lr;
        if (R.l + 2 <= w.n_refc || !huge) lr = warp_local_align(R, R.l, Q, len, w);
        else {
            // a scanned window: the slices found where the forward pass ends; everything else happens in a narrow
            // sub-window that ends at that column (the reverse pass and the final global alignment never leave it)
            int jid = -1;
            for (int t = 0; t < huge->n_jobs; ++t)
                if (huge->jobs[t].pair == pair && huge->jobs[t].beg == (long long)*beg && huge->jobs[t].reglen == reglen && huge->jobs[t].strand == Q.strand) { jid = t; break; }
            if (jid < 0) return -1;
            const HugeJob &J = huge->jobs[jid];
            ScanBest best; best.score = 0; best.j = 0; best.i = 0;
            for (int t = 0; t < J.n_slices; ++t) {
                const ScanBest sb = huge->slice_best[J.first_slice + t];
                if (sb.score < 0) return -1;
                if (sb.score > best.score || (sb.score == best.score && sb.score > 0 && (sb.j < best.j || (sb.j == best.j && sb.i < best.i)))) best = sb;
            }
            lr.score = best.score; lr.n_ops = 0; lr.start_i = lr.start_j = lr.end_i = lr.end_j = 0; lr.too_big = false;
            if (best.score >= 1) {
                const int span = w.n_refc - 2 < 640 ? w.n_refc - 2 : 640;
                const int off = best.i > span ? best.i - span : 0;
                RefWin R2; R2.pac = pac; R2.beg = R.beg + off; R2.l = best.i - off;
                lr = warp_local_align(R2, R2.l, Q, len, w);
                if (lr.too_big || lr.end_i != R2.l || lr.end_j != best.j) return -1;       // must end where the full pass ends
                lr.start_i += off; lr.end_i += off;
            }
        }
        if (lr.too_big) return -1;
        int nc = 0;
        long long b = *beg;
        uint32_t c = 0, c0 = 0, cl = 0;
        if (lane == 0) {
            DpScratch sc; sc.ints = nullptr; sc.n_ints = 0; sc.bytes = w.ops; sc.n_bytes = w.n_ops; sc.istride = sc.bstride = 1;
            int64_t bb = *beg;
            nc = sw_post(R, Q, lr, &bb, cigar, &c, sc);
            b = bb;
            if (nc > 0) { c0 = cigar[0]; cl = cigar[nc - 1]; }
        }
        nc = __shfl_sync(FQB_FULL, nc, 0);
        b = __shfl_sync(FQB_FULL, b, 0);
        c = __shfl_sync(FQB_FULL, c, 0);
        c0 = __shfl_sync(FQB_FULL, c0, 0);
        cl = __shfl_sync(FQB_FULL, cl, 0);
        if (nc > 0 && lane) { cigar[0] = (uint16_t)c0; cigar[nc - 1] = (uint16_t)cl; }
        *beg = b; *cnt = c;
        return nc;
    }
};

// refine_gapped_core by a warp; the CIGAR and position are valid in lane 0, the return value in all lanes
__device__ int warp_refine_gapped(int64_t l_pac, const uint8_t *pac, const ReadSeq &Q, uint32_t *pos_io, int ext, uint16_t *cigar, int cap,
                                  const WarpDp &w) {
    int64_t pos;
    const RefWin R = refine_window(l_pac, pac, Q.len, *pos_io, ext, &pos);
    if (!warp_load_ref(R, w)) return -1;
    GlobalResult g = warp_global_any(R, 0, R.l, Q, 0, Q.len, kGapEnd, kBandWidth, w);
    if (g.too_big) return -1;
    int nc = 0;
    if (w.lane == 0) {
        DpScratch sc; sc.ints = nullptr; sc.n_ints = 0; sc.bytes = w.ops; sc.n_bytes = w.n_ops; sc.istride = sc.bstride = 1;
        nc = refine_post(g, pos, ext, pos_io, cigar, cap, sc);
    }
    return __shfl_sync(FQB_FULL, nc, 0);
}

}  // namespace fqb
