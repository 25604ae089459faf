// Wavefront form of the dynamic-programming passes of rows a10/a11 (one alignment per warp, device only).
//
// Each lane owns CT CONSECUTIVE columns of the DP matrix and keeps the previous row of those columns in REGISTERS; lane l
// works on row t - l + 1 at step t, so the only traffic between lanes is what column i0 - 1 hands to column i0: one
// __shfl_up of two or three values per step.  Inside a lane the cells of a row are computed in column order with exactly the
// recurrences, boundary rules and tie-breaks of the per-lane statement in fq_device_dp.cuh (aln_local_core's forward
// pass, libbwa/stdaln.c:529-640; aln_global_core, stdaln.c:345-524) -- there is no prefix scan and no shared-memory row,
// which is where the row-chunk form of fq_dp_warp.cuh spent its instructions (about 100 per 32 cells; here 15-30).
// The trace-back matrix of the banded global alignment is four bits per cell, one word per lane and row, in shared memory.
#pragma once
#include "fq_device_dp.cuh"

namespace fqb {

#ifndef FQB_FULL
#define FQB_FULL 0xffffffffu
#endif

// max(a + b, c) and max(a, b, c): single DPX instructions (VIADDMNMX / VIMNMX3) on sm_90+
__device__ __forceinline__ int dp_addmax(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    const int s = a + b; return s > c ? s : c;
#endif
}
__device__ __forceinline__ int dp_max3(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_s32(a, b, c);
#else
    const int m = a > b ? a : b; return m > c ? m : c;
#endif
}

// reference codes of a lane's CT columns, four bits each (8 = a padding column that matches no read base)
template <int CT> struct LaneCodes {
    static constexpr int NW = (CT + 7) / 8;
    uint32_t w[NW];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = 0;
    }
    __device__ __forceinline__ void set(int c, uint32_t code) { w[c >> 3] |= code << (4 * (c & 7)); }
};

// ---- aln_local_core, forward pass: best score and the first cell (row-major) that reaches it ----
// refc[i - 1] = code of column i (1-based), len1 <= 32 * CT.  All lanes return the same values.
template <int CT>
__device__ __forceinline__ void wave_local_forward(const uint8_t *refc, int len1, const ReadSeq &Q, int len2, int lane, int &score_f, int &end_i, int &end_j) {
    constexpr int NW = LaneCodes<CT>::NW;
    const int r = kGapExt, qr = kGapOpen + kGapExt;
    const int i0 = lane * CT;                              // this lane's columns are i0 + 1 .. i0 + CT
    LaneCodes<CT> cw; cw.clear();
#pragma unroll
    for (int c = 0; c < CT; ++c) cw.set(c, i0 + c < len1 ? (uint32_t)refc[i0 + c] : 8u);
    int H[CT], E[CT];                                      // h and e of the previous row
#pragma unroll
    for (int c = 0; c < CT; ++c) { H[c] = 0; E[c] = 0; }
    int oh = 0, of = 0;                                    // h and f after the last column of the row this lane finished last
    int ph = 0;                                            // h of column i0 in the previous row (the diagonal of column i0 + 1)
    int best = 0, bpos = 0;                                // first maximum in this lane's own row-major order; bpos = j << 16 | i
    const int steps = len2 + 31;
    for (int t = 0; t < steps; ++t) {
        int ih = __shfl_up_sync(FQB_FULL, oh, 1), iff = __shfl_up_sync(FQB_FULL, of, 1);
        if (lane == 0) { ih = 0; iff = 0; }                // column 0: last_h = 0, f = 0 at the start of a row
        const int j = t - lane + 1;
        if (j >= 1 && j <= len2) {
            const uint32_t qj = Q.at(j - 1);
            const bool qn = qj > 3;
            const int s_eq = qn ? -13 : 11, s_ne = qn ? -13 : -19;      // aln_sm_maq, the reference base is never N
            uint32_t x[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) x[k] = cw.w[k] ^ ((qj & 3u) * 0x11111111u);
            int dg = ph, lh = ih, lf = iff;
            ph = ih;
            const int jpos = j << 16 | (i0 + 1);
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int hu = H[c], eu = E[c];
                const int sc = (x[c >> 3] & (15u << (4 * (c & 7)))) ? s_ne : s_eq;
                int h = dp_addmax(dg, sc, 0);
                // f = (f > last_h - q ? f : last_h - q) - r.  The reference skips the update while last_h == 0; then f <= 0
                // already and stays <= 0 either way, and a non-positive f never changes a cell (h >= 0)
                lf = dp_addmax(lf, -r, lh - qr);
                const int e = hu > qr ? dp_addmax(eu, -r, hu - qr) : 0;
                h = dp_max3(h, lf, e);
                H[c] = h; E[c] = e; dg = hu; lh = h;
                if (h > best) { best = h; bpos = jpos + c; }
            }
            oh = lh; of = lf;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {                     // the maximum; among equals the earliest row, then the smallest column
        const int ob = __shfl_xor_sync(FQB_FULL, best, d), op = __shfl_xor_sync(FQB_FULL, bpos, d);
        if (ob > best || (ob == best && op < bpos)) { best = ob; bpos = op; }
    }
    score_f = best; end_j = bpos >> 16; end_i = bpos & 0xffff;
}

// ---- aln_global_core (banded) ----
// Columns 1 .. len1 over the lanes (lane l owns l CT + 1 .. l CT + CT, len1 <= 32 CT); column 0 and row 0 are closed forms
// (M(0,0) = 0, the leading end-gap chains) that lane 0 / the first step substitute for what a neighbour would have sent.
// The band is applied by masking: a cell outside (first, endc] of its row holds MINOR_INF in all three states, which is what the
// reference's boundary cell holds and what its "no cell above the last column" rule amounts to (that rule itself is kept
// explicitly so that even the never-used values agree).  Per row, the band limits, the end-gap switches and the read base come
// from one word of `rowinfo` (shared memory, filled once per alignment).
// trace[(j - 1) * 32 + lane] holds the cells (j, l CT + 1 .. l CT + CT), four bits each: bit 0 the M cell came from M; else bit 1
// says I (set) or D; bit 2 set = the I cell extends an insertion, bit 3 set = the D cell extends a deletion (clear = opened from M).
// The path's ctype sequence is left in ops[0 .. n_ops), from the END of the alignment to its start.
template <int CT, typename TW>
__device__ __forceinline__ GlobalResult wave_global_align(const uint8_t *refc, int r0, int len1, const ReadSeq &Q, int q0, int len2, int gap_end, int band,
                                                          int32_t *rowinfo, TW *trace, int trace_rows, uint8_t *ops, int ops_cap, int lane) {
    static_assert(CT <= 8 && CT * 4 <= (int)sizeof(TW) * 8, "one trace word per lane and row");
    GlobalResult res; res.score = 0; res.n_ops = 0; res.too_big = false;
    if (len1 == 0 || len2 == 0) return res;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; } else { b1 = band; b2 = len2 - len1 + band; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const int end_ge = gap_end >= 0 ? gap_end : kGapExt;
    if (len1 > 32 * CT || len1 > 510 || len2 > trace_rows || len1 + len2 + 2 > ops_cap) { res.too_big = true; return res; }
    {   // per-row constants: first | endc << 9 | head << 18 | last_have_up << 19 | (last_ige is the end-gap one) << 20 | (d_ge is) << 21 | read base << 24
        const int tmp_end = (b2 < len2) ? b2 : len2 - 1;
        for (int j = 1 + lane; j <= len2; j += 32) {
            const bool head = j <= tmp_end || (j == tmp_end + 1 && j == len2 && b2 != len2 - 1);   // band starts at column 0
            const bool mid = !head && j <= len2 - b2 + 1;                                         // right edge inside the matrix
            const bool last_row_d = head ? (j == tmp_end + 1) : (!mid && j == len2);              // set_end_D rows
            int first, endc;
            if (head) { first = 0; endc = (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1; }
            else { first = j - b2; endc = mid ? j + b1 - 1 : len1; }
            const bool last_have_up = head ? (j + b1 - 1 > len1) : !mid;                          // the cell above the last column exists
            const bool last_ige_end = head || !mid;                                                // set_end_I on the last column
            rowinfo[j] = first | endc << 9 | (int)head << 18 | (int)last_have_up << 19 | (int)last_ige_end << 20 | (int)last_row_d << 21 | (int)Q.at(q0 + j - 1) << 24;
        }
        __syncwarp();
    }
    const int i0 = lane * CT;                              // this lane's columns are i0 + 1 .. i0 + CT
    LaneCodes<CT> cw; cw.clear();
#pragma unroll
    for (int c = 0; c < CT; ++c) cw.set(c, i0 + c < len1 ? (uint32_t)refc[r0 + i0 + c] : 8u);
    // row 0: the D chain with the end-gap extension over columns 1 .. b1 - 1
    int pM[CT], pI[CT], pD[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) { const int i = i0 + c + 1; pM[c] = kNegInf; pI[c] = kNegInf; pD[c] = i < b1 ? -(kGapOpen + i * end_ge) : kNegInf; }
    // row-0 values of column i0 (the diagonal of this lane's first column in row 1)
    int sM = i0 == 0 ? 0 : kNegInf, sI = kNegInf, sD = (i0 >= 1 && i0 < b1) ? -(kGapOpen + i0 * end_ge) : kNegInf;
    const int n_lanes = (len1 + CT - 1) / CT;
    const int steps = len2 + n_lanes - 1;
    for (int t = 0; t < steps; ++t) {
        int rM = __shfl_up_sync(FQB_FULL, pM[CT - 1], 1), rI = __shfl_up_sync(FQB_FULL, pI[CT - 1], 1), rD = __shfl_up_sync(FQB_FULL, pD[CT - 1], 1);
        const int j = t - lane + 1;
        if (j >= 1 && j <= len2) {
            if (lane == 0) {                               // column 0 of row j: M = D = MINOR_INF, I = the leading end-gap chain (head rows; masked otherwise)
                rM = kNegInf; rD = kNegInf; rI = -(kGapOpen + j * end_ge);
            }
            const int info = rowinfo[j];
            const int lo = (info & 511) - i0, hi = ((info >> 9) & 511) - i0;     // cells lo <= c < hi are inside the band
            const bool have_up = info & (1 << 19);
            const int last_ige = (info & (1 << 20)) ? end_ge : kGapExt;
            const int d_ge = (info & (1 << 21)) ? end_ge : kGapExt;
            const uint32_t qj = (uint32_t)info >> 24;
            const bool qn = qj > 3;
            const int s_eq = qn ? -13 : 11, s_ne = qn ? -13 : -19;
            const uint32_t x = cw.w[0] ^ ((qj & 3u) * 0x11111111u);
            int lM = rM, lD = rD;                          // (j, i0)
            int dM = sM, dI = sI, dD = sD;                 // (j - 1, i0)
            sM = rM; sI = rI; sD = rD;
            const int go_d = kGapOpen + d_ge;
            uint32_t tw = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const bool inb = c >= lo && c < hi, lastc = c == hi - 1;
                const int uM = pM[c], uI = pI[c], uD = pD[c];
                const int sco = (x & (15u << (4 * c))) ? s_ne : s_eq;
                const int mx = dp_max3(dM, dI, dD);
                int m = mx + sco;
                if (dM == mx) tw |= 1u << (4 * c);
                if (dI > dD) tw |= 2u << (4 * c);
                const int ige = lastc ? last_ige : kGapExt;
                const int ai = uM - kGapOpen - ige, bi = uI - ige;
                int iv = ai > bi ? ai : bi;
                const bool noup = lastc && !have_up;
                if (!(ai > bi) && !noup) tw |= 4u << (4 * c);
                if (noup) iv = kNegInf;
                const int ad = lM - go_d, bd = lD - d_ge;
                int d = ad > bd ? ad : bd;
                if (!(ad > bd)) tw |= 8u << (4 * c);
                if (!inb) { m = kNegInf; iv = kNegInf; d = kNegInf; }
                pM[c] = m; pI[c] = iv; pD[c] = d;
                lM = m; lD = d; dM = uM; dI = uI; dD = uD;
            }
            trace[(j - 1) * 32 + lane] = (TW)tw;
        }
    }
    __syncwarp();
    // the three scores of cell (len2, len1)
    int fM = 0, fI = 0, fD = 0;
    {
        const int ce = (len1 - 1) % CT;
#pragma unroll
        for (int c = 0; c < CT; ++c) if (c == ce) { fM = pM[c]; fI = pI[c]; fD = pD[c]; }
        const int le = (len1 - 1) / CT;
        fM = __shfl_sync(FQB_FULL, fM, le); fI = __shfl_sync(FQB_FULL, fI, le); fD = __shfl_sync(FQB_FULL, fD, le);
    }
    int score = 0, n_ops = 0;
    if (lane == 0) {                                       // back-trace (stdaln.c:480-512)
        // cell (j, i): bits as above; row 0 is the leading deletion chain, column 0 the leading insertion chain
        auto cell_at = [&](int j, int i) -> uint32_t {
            if (j == 0) return i <= 1 ? 0u : 8u;
            if (i == 0) return j == 1 ? 0u : 4u;
            return ((uint32_t)trace[(j - 1) * 32 + (i - 1) / CT] >> (4 * ((i - 1) % CT))) & 15u;
        };
        auto m_src = [](uint32_t cell) -> uint32_t { return (cell & 1u) ? kOpM : ((cell & 2u) ? kOpI : kOpD); };
        int i = len1, j = len2;
        int mx = fM;
        uint32_t cell = cell_at(j, i);
        uint32_t type = m_src(cell), ctype = kOpM;
        if (fI > mx) { mx = fI; type = (cell & 4u) ? kOpI : kOpM; ctype = kOpI; }
        if (fD > mx) { mx = fD; type = (cell & 8u) ? kOpD : kOpM; ctype = kOpD; }
        int n = 0;
        ops[n++] = (uint8_t)ctype;
        do {
            if (ctype == kOpM) { --i; --j; } else if (ctype == kOpI) --j; else --i;
            cell = cell_at(j, i);
            ctype = type;
            type = ctype == kOpM ? m_src(cell) : ctype == kOpI ? ((cell & 4u) ? kOpI : kOpM) : ((cell & 8u) ? kOpD : kOpM);
            ops[n++] = (uint8_t)ctype;
        } while (i || j);
        score = mx; n_ops = n - 1;
    }
    res.score = __shfl_sync(FQB_FULL, score, 0);
    res.n_ops = __shfl_sync(FQB_FULL, n_ops, 0);
    __syncwarp();
    return res;
}

}  // namespace fqb
