// Per-thread device logic of the dynamic-programming rows (a10, a11): banded global
// alignment (aln_global_core, libbwa/stdaln.c:345-524), local alignment
// (aln_local_core, stdaln.c:529-757), bwa_sw_core / bwa_paired_sw (libbwa/bwape.c:359-625)
// and bwa_refine_gapped with NM and bwa_correct_trimmed (libbwa/bwase.c:183-418).
//
// One alignment per thread.  All DP state lives in a per-thread scratch area that is
// INTERLEAVED across the 32 lanes of a warp (element e of lane l at e*stride + l), so the
// lanes of a warp, which walk their matrices in lock step, issue fully coalesced accesses.
// On the host (tests/emul) stride is 1.
#pragma once
#include "fq_device_pair.cuh"

namespace fqb {

constexpr int kNegInf = -1073741823;          // MINOR_INF, libbwa/stdaln.h:97
constexpr int kOpM = 0, kOpI = 1, kOpD = 2, kOpS = 3;
constexpr int kGapOpen = 26, kGapExt = 9, kGapEnd = 5, kBandWidth = 50;   // aln_param_bwa, stdaln.c:231

#if defined(__CUDA_ARCH__)
#define FQB_DADD(a, b) __dadd_rn((a), (b))
#define FQB_DSUB(a, b) __dsub_rn((a), (b))
#define FQB_DMUL(a, b) __dmul_rn((a), (b))
#else
#define FQB_DADD(a, b) ((a) + (b))
#define FQB_DSUB(a, b) ((a) - (b))
#define FQB_DMUL(a, b) ((a) * (b))
#endif

// aln_sm_maq (stdaln.c:206-212)
FQB_HD int maq_score(uint32_t a, uint32_t b) { return (a > 3 || b > 3) ? -13 : (a == b ? 11 : -19); }

struct DpScratch {
    int32_t *ints; int n_ints;       // interleaved int scratch (rows of the DP): shared memory in the fast kernels
    uint8_t *bytes; int n_bytes;     // interleaved byte scratch (trace-back matrix, ops): global memory
    int istride, bstride;            // element e of this lane at e * stride (threads per block / 32 on the device, 1 on the host)
    FQB_HD int32_t &I(int e) const { return ints[(size_t)e * istride]; }
    FQB_HD uint8_t &B(int e) const { return bytes[(size_t)e * bstride]; }
};

// reference window and read, both as "sequence of nt4 codes" accessors
struct RefWin {                      // l bases of the packed reference starting at pac coordinate beg
    const uint8_t *pac; int64_t beg; int l;
    FQB_HD uint32_t at(int i) const { int64_t k = beg + i; return pac[k >> 2] >> ((~k & 3) << 1) & 3; }   // i is 0-based
};
struct ReadSeq {                     // the read in alignment orientation (strand 1 = reverse complement)
    const uint8_t *fwd; int len; int strand;
    FQB_HD uint32_t at(int j) const {
        uint32_t c = strand ? fwd[len - 1 - j] : fwd[j];
        return (strand && c < 4) ? 3 - c : c;
    }
};

struct GlobalResult { int score; int n_ops; bool too_big; };

// set_M / set_I / set_D families (stdaln.c:260-318)
FQB_HD void dp_from_diag(int pM, int pI, int pD, int sc, int &out, uint32_t &t) {
    if (pM >= pI) { if (pM >= pD) { out = pM + sc; t = kOpM; } else { out = pD + sc; t = kOpD; } }
    else { if (pI > pD) { out = pI + sc; t = kOpI; } else { out = pD + sc; t = kOpD; } }
}
FQB_HD void dp_gap(int pM, int pG, int go, int ge, uint32_t self, int &out, uint32_t &t) {
    if (pM - go > pG) { t = kOpM; out = pM - go - ge; } else { t = self; out = pG - ge; }
}

// Banded global alignment of ref window columns [r0, r0+len1) against read rows [q0, q0+len2).
// The path's ctype sequence is left in sc.B(ops_base + k), k = 0..n_ops-1, from the END of the
// alignment to its start (= path[] of the reference).  Trace cells: mt | it<<2 | dt<<4.
FQB_HD GlobalResult global_align(const RefWin &R, int r0, int len1, const ReadSeq &Q, int q0, int len2, int gap_end, int band,
                                 const DpScratch &sc, int ops_base) {
    GlobalResult res; res.score = 0; res.n_ops = 0; res.too_big = false;
    if (len1 == 0 || len2 == 0) return res;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; } else { b1 = band; b2 = len2 - len1 + band; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const int W = len1 + 1;
    const int tw = (b1 + b2 <= len1) ? (b1 + b2 + 1) : W;          // trace columns kept per row
    const int end_ge = gap_end >= 0 ? gap_end : kGapExt;
    if (6 * W > sc.n_ints || (len2 + 1) * tw + ops_base + len1 + len2 + 2 > sc.n_bytes) { res.too_big = true; return res; }
    const int tr_base = ops_base + len1 + len2 + 2;
    // rows: cur/last x {M,I,D}; cell (row, i) at ints[(row*3 + which) * W + i]
#define ROW(rw, which, i) sc.I(((rw) * 3 + (which)) * W + (i))
#define TRC(j, i) sc.B(tr_base + (j) * tw + ((i) - ((j) > b2 ? (j) - b2 : 0)))
    int cur = 0, last = 1;
    ROW(cur, 0, 0) = 0; ROW(cur, 1, 0) = kNegInf; ROW(cur, 2, 0) = kNegInf;
    for (int i = 1; i < b1; ++i) {
        int d; uint32_t t;
        dp_gap(ROW(cur, 0, i - 1), ROW(cur, 2, i - 1), kGapOpen, end_ge, kOpD, d, t);
        ROW(cur, 0, i) = kNegInf; ROW(cur, 1, i) = kNegInf; ROW(cur, 2, i) = d;
        TRC(0, i) = (uint8_t)(t << 4);
    }
    { int t = cur; cur = last; last = t; }
    const int tmp_end = (b2 < len2) ? b2 : len2 - 1;
    for (int j = 1; j <= len2; ++j) {
        const bool head = j <= tmp_end || (j == tmp_end + 1 && j == len2 && b2 != len2 - 1);   // band starts at column 0
        const bool mid = !head && j <= len2 - b2 + 1;                                         // right edge inside the matrix
        const bool last_row_d = head ? (j == tmp_end + 1) : (!mid && j == len2);              // set_end_D rows
        const int d_ge = last_row_d ? end_ge : kGapExt;
        const uint32_t qj = Q.at(q0 + j - 1);
        int first, endc;
        if (head) {
            first = 0;
            endc = (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1;
            int iv; uint32_t t;
            dp_gap(ROW(last, 0, 0), ROW(last, 1, 0), kGapOpen, end_ge, kOpI, iv, t);
            ROW(cur, 0, 0) = kNegInf; ROW(cur, 1, 0) = iv; ROW(cur, 2, 0) = kNegInf;
            TRC(j, 0) = (uint8_t)(t << 2);
        } else {
            first = j - b2;
            endc = mid ? j + b1 - 1 : len1;
            ROW(cur, 0, first) = kNegInf; ROW(cur, 1, first) = kNegInf; ROW(cur, 2, first) = kNegInf;
        }
        int lM = ROW(cur, 0, first), lD = ROW(cur, 2, first);          // left neighbour in this row
        int dM = ROW(last, 0, first), dI = ROW(last, 1, first), dD = ROW(last, 2, first);   // diagonal neighbour
        for (int i = first + 1; i <= endc; ++i) {
            const int sco = maq_score(R.at(r0 + i - 1), qj);
            int m, iv, d; uint32_t tm, ti = 0, td;
            dp_from_diag(dM, dI, dD, sco, m, tm);
            const bool lastc = i == endc;
            int uM = 0, uI = 0, uD = 0;
            bool have_up = true;
            if (lastc) {
                if (head) have_up = j + b1 - 1 > len1;
                else if (mid) have_up = false;
            }
            if (have_up) {
                uM = ROW(last, 0, i); uI = ROW(last, 1, i); uD = ROW(last, 2, i);
                const int ige = (lastc && (head || !mid)) ? end_ge : kGapExt;   // set_end_I on the last column
                dp_gap(uM, uI, kGapOpen, ige, kOpI, iv, ti);
            } else iv = kNegInf;
            dp_gap(lM, lD, kGapOpen, d_ge, kOpD, d, td);
            ROW(cur, 0, i) = m; ROW(cur, 1, i) = iv; ROW(cur, 2, i) = d;
            TRC(j, i) = (uint8_t)(tm | ti << 2 | td << 4);
            lM = m; lD = d;
            dM = uM; dI = uI; dD = uD;
        }
        { int t = cur; cur = last; last = t; }
    }
    // backtrace
    int i = len1, j = len2;
    int mx = ROW(last, 0, len1);
    uint32_t cell = TRC(j, i), type = cell & 3, ctype = kOpM;
    if (ROW(last, 1, len1) > mx) { mx = ROW(last, 1, len1); type = (cell >> 2) & 3; ctype = kOpI; }
    if (ROW(last, 2, len1) > mx) { mx = ROW(last, 2, len1); type = (cell >> 4) & 3; ctype = kOpD; }
    int n = 0;
    sc.B(ops_base + n++) = (uint8_t)ctype;
    do {
        if (ctype == kOpM) { --i; --j; } else if (ctype == kOpI) --j; else --i;
        cell = (i == 0 && j == 0) ? 0 : TRC(j, i);
        ctype = type;
        type = ctype == kOpM ? (cell & 3) : ctype == kOpI ? ((cell >> 2) & 3) : ((cell >> 4) & 3);
        sc.B(ops_base + n++) = (uint8_t)ctype;
    } while (i || j);
    res.score = mx; res.n_ops = n - 1;
#undef ROW
#undef TRC
    return res;
}

// aln_path2cigar32 + bwa_aln_path2cigar: runs from the start of the alignment; returns n_cigar (or -1 if > cap)
FQB_HD int ops_to_cigar(const DpScratch &sc, int ops_base, int n_ops, uint16_t *cigar, int cap) {
    if (n_ops == 0) return 0;
    int n = 0;
    cigar[0] = (uint16_t)(sc.B(ops_base + n_ops - 1) << 14 | 1);
    for (int i = n_ops - 2; i >= 0; --i) {
        uint32_t op = sc.B(ops_base + i);
        if (op == (uint32_t)(cigar[n] >> 14)) cigar[n] += 1;
        else { if (n + 1 >= cap) return -1; cigar[++n] = (uint16_t)(op << 14 | 1); }
    }
    return n + 1;
}

// refine_gapped_core with is_end_correct = 1 (libbwa/bwase.c:183-232), in three pieces so that the warp-cooperative
// kernel can share the scalar parts: the reference window, the alignment, and the CIGAR/position fix-up.
FQB_HD RefWin refine_window(int64_t l_pac, const uint8_t *pac, int len, uint32_t pos_in, int ext, int64_t *pos_out) {
    const int ref_len = len + (ext < 0 ? -ext : ext);
    int64_t pos = pos_in > (uint32_t)l_pac ? (int64_t)(int32_t)pos_in : (int64_t)pos_in;
    RefWin R; R.pac = pac;
    if (ext > 0) { R.beg = pos; int64_t e = pos + ref_len < l_pac ? pos + ref_len : l_pac; R.l = (int)(e > pos ? e - pos : 0); }
    else {
        int64_t x = pos + len, b = x - ref_len > 0 ? x - ref_len : 0, e = x < l_pac ? x : l_pac;
        R.beg = b; R.l = (int)(e > b ? e - b : 0);
    }
    *pos_out = pos;
    return R;
}
// path ops at sc.B(0 ..); returns n_cigar (-1 = capacity) and the corrected position
FQB_HD int refine_post(const GlobalResult &g, int64_t pos, int ext, uint32_t *pos_io, uint16_t *cigar, int cap, const DpScratch &sc) {
    int n_cigar = ops_to_cigar(sc, 0, g.n_ops, cigar, cap);
    if (n_cigar <= 0) return n_cigar < 0 ? -1 : 0;
    if (ext < 0) {
        int d = 0;
        for (int k = 0; k < n_cigar; ++k) {
            if ((cigar[k] >> 14) == kOpD) d -= cigar[k] & 0x3fff;
            else if ((cigar[k] >> 14) == kOpI) d += cigar[k] & 0x3fff;
        }
        pos += d;
    }
    if ((cigar[0] >> 14) == kOpD) {
        pos += cigar[0] & 0x3fff;
        for (int k = 0; k < n_cigar - 1; ++k) cigar[k] = cigar[k + 1];
        --n_cigar;
    }
    if ((cigar[n_cigar - 1] >> 14) == kOpD) --n_cigar;
    if ((cigar[n_cigar - 1] >> 14) == kOpI) cigar[n_cigar - 1] = (uint16_t)(kOpS << 14 | (cigar[n_cigar - 1] & 0x3fff));
    if ((cigar[0] >> 14) == kOpI) cigar[0] = (uint16_t)(kOpS << 14 | (cigar[0] & 0x3fff));
    *pos_io = (uint32_t)pos;
    return n_cigar;
}
FQB_HD int refine_gapped(int64_t l_pac, const uint8_t *pac, const ReadSeq &Q, uint32_t *pos_io, int ext, uint16_t *cigar, int cap,
                         const DpScratch &sc) {
    int64_t pos;
    const RefWin R = refine_window(l_pac, pac, Q.len, *pos_io, ext, &pos);
    GlobalResult g = global_align(R, 0, R.l, Q, 0, Q.len, kGapEnd, kBandWidth, sc, 0);
    if (g.too_big) return -1;
    return refine_post(g, pos, ext, pos_io, cigar, cap, sc);
}

// NM exactly as bwa_cal_md1 counts it (libbwa/bwase.c:234-296)
FQB_HD int cal_nm(const fqb_read_t &s, const ReadSeq &Q, int64_t l_pac, const uint8_t *pac) {
    int nm = 0;
    uint32_t x = s.pos, y = 0;
    if (s.has_cigar) {
        for (int k = 0; k < s.n_cigar; ++k) {
            const int l = s.cigar[k] & 0x3fff, op = s.cigar[k] >> 14;
            if (op == kOpM) {
                for (int z = 0; z < l && (int64_t)x + z < l_pac; ++z) {
                    const int64_t kk = (int64_t)x + z;
                    const uint32_t c = pac[kk >> 2] >> ((~kk & 3) << 1) & 3, q = Q.at((int)(y + z));
                    if (q > 3 || c != q) ++nm;
                }
                x += l; y += l;
            } else if (op == kOpI || op == kOpS) { y += l; if (op == kOpI) nm += l; }
            else { x += l; nm += l; }
        }
    } else {
        for (int z = 0; z < s.len; ++z) {
            int64_t k = (int64_t)x + z;
            uint32_t c = pac[k >> 2] >> ((~k & 3) << 1) & 3, q = Q.at(z);
            if (q > 3 || c != q) ++nm;
        }
    }
    return nm;
}

// bwa_correct_trimmed (libbwa/bwase.c:298-337)
FQB_HD void correct_trimmed(fqb_read_t &s) {
    const int clip = s.full_len - s.len;
    if (clip == 0) return;
    if (s.strand == 0) {
        if (s.has_cigar && (s.cigar[s.n_cigar - 1] >> 14) == kOpS) s.cigar[s.n_cigar - 1] += (uint16_t)clip;
        else {
            if (!s.has_cigar) { s.n_cigar = 2; s.has_cigar = 1; s.cigar[0] = (uint16_t)(kOpM << 14 | s.len); }
            else ++s.n_cigar;
            s.cigar[s.n_cigar - 1] = (uint16_t)(kOpS << 14 | clip);
        }
    } else {
        if (s.has_cigar && (s.cigar[0] >> 14) == kOpS) s.cigar[0] += (uint16_t)clip;
        else {
            if (!s.has_cigar) { s.n_cigar = 2; s.has_cigar = 1; s.cigar[1] = (uint16_t)(kOpM << 14 | s.len); }
            else { ++s.n_cigar; for (int k = s.n_cigar - 1; k > 0; --k) s.cigar[k] = s.cigar[k - 1]; }
            s.cigar[0] = (uint16_t)(kOpS << 14 | clip);
        }
    }
    s.len = s.full_len;
}

// ---- local alignment (aln_local_core, _thres = 1, no _subo).  Scores stay below the reference's
// overflow threshold (32000) for reads <= 256 bp, so its rescaling blocks cannot run.
struct LocalResult { int score; int n_ops; int start_i, start_j, end_i, end_j; bool too_big; };

FQB_HD LocalResult local_align(const RefWin &R, int len1, const ReadSeq &Q, int len2, const DpScratch &sc, int ops_base) {
    LocalResult res; res.score = -1; res.n_ops = 0; res.start_i = res.start_j = res.end_i = res.end_j = 0; res.too_big = false;
    if (len1 == 0 || len2 == 0) return res;
    const int q = kGapOpen, r = kGapExt, qr = q + r, max_score = 11;
    if (len1 + 2 > sc.n_ints) { res.too_big = true; return res; }
    // one word per column, h in the high half and e in the low half, as the reference packs eh[] (stdaln.c:252-256)
#define EH(i) (sc.I(i) >> 16)
#define EE(i) (sc.I(i) & 0xffff)
#define EHE_SET(i, h, e) (sc.I(i) = ((h) << 16) | (e))
    for (int i = 0; i <= len1 + 1; ++i) EHE_SET(i, 0, 0);
    int score_f = 0, end_i = 0, end_j = 0;
    for (int j = 1; j <= len2; ++j) {
        int last_h = 0, f = 0;
        const uint32_t qj = Q.at(j - 1);
        int h_here = EH(0), e_here = EE(0);
        for (int i = 1; i <= len1; ++i) {
            const int w_next = sc.I(i), h_next = w_next >> 16, e_next = w_next & 0xffff;
            int curr_h = h_here + maq_score(R.at(i - 1), qj);
            if (curr_h < 0) curr_h = 0;
            if (last_h > 0) { f = (f > last_h - q) ? f - r : last_h - qr; if (curr_h < f) curr_h = f; }
            int e = 0;
            if (h_next >= qr + 1) {
                e = (e_here > h_next - q) ? e_here - r : h_next - qr;
                if (curr_h < e) curr_h = e;
            }
            EHE_SET(i - 1, last_h, e);
            last_h = curr_h;
            if (score_f < curr_h) { score_f = curr_h; end_i = i; end_j = j; }
            h_here = h_next; e_here = e_next;
        }
        EHE_SET(len1, last_h, 0);
    }
    res.score = score_f;
    if (score_f < 1) return res;
    for (int i = end_i; i >= 0; --i) EHE_SET(i, 0, 0);
    if (end_i == 0 || end_j == 0) return res;
    int score_r = maq_score(R.at(end_i - 1), Q.at(end_j - 1));
    int start_i = end_i, start_j = end_j;
    EHE_SET(end_i, qr + score_r, 0);
    int start = end_i - 1, end = end_i - 3;
    if (end <= 0) end = 0;
    for (int j = end_j - 1; j != 0; --j) {
        int last_h = 0, f = 0;
        bool stop = false;
        const uint32_t qj = Q.at(j - 1);
        int i = start;
        for (; i != end; --i) {
            const int w_here = sc.I(i + 1);
            int curr_h = (w_here >> 16) + maq_score(R.at(i - 1), qj);
            if (curr_h < 0) curr_h = 0;
            if (last_h > 0) { f = (f > last_h - q) ? f - r : last_h - qr; if (curr_h < f) curr_h = f; }
            const int curr_last_h = EH(i), e_old = w_here & 0xffff;
            int e = (e_old > curr_last_h - q) ? e_old - r : curr_last_h - qr;
            if (e < 0) e = 0;
            if (curr_h < e) curr_h = e;
            EHE_SET(i + 1, last_h, e);
            last_h = curr_h;
            if (score_r < curr_h) {
                score_r = curr_h; start_i = i; start_j = j;
                if (score_r - qr == score_f) { stop = true; break; }
            }
        }
        if (stop) break;
        EHE_SET(i + 1, last_h, 0);
        if (EH(start) <= qr) --start;
        if (start <= 0) start = 0;
        end = start_i - (start_j - j) - (score_r + (start_j - j) * max_score) / r - 1;
        if (end <= 0) end = 0;
    }
#undef EH
#undef EE
#undef EHE_SET
    score_r -= qr;
    int jmax = (end_i - start_i > end_j - start_j) ? end_i - start_i : end_j - start_j;
    ++jmax;
    GlobalResult g;
    for (int w = kBandWidth;; w <<= 1) {
        g = global_align(R, start_i - 1, end_i - start_i + 1, Q, start_j - 1, end_j - start_j + 1, -1, w, sc, ops_base);
        if (g.too_big) { res.too_big = true; return res; }
        if (g.score == score_r || score_f == g.score) break;
        if (w > jmax) break;
    }
    res.score = (score_r > g.score && score_f > g.score) ? -1 : g.score;
    res.n_ops = g.n_ops;
    res.start_i = start_i; res.start_j = start_j; res.end_i = end_i; res.end_j = end_j;
    return res;
}

constexpr int kSwCigarCap = 48;

// tail of bwa_sw_core (libbwa/bwape.c:385-445): CIGAR, clipping and mismatch/gap counts from the local alignment whose
// path ops sit at sc.B(0 ..).  Returns n_cigar (0 = none), -1 = cigar capacity exceeded.
FQB_HD int sw_post(const RefWin &R, const ReadSeq &Q, const LocalResult &lr, int64_t *beg, uint16_t *cigar, uint32_t *cnt, const DpScratch &sc) {
    const int len = Q.len;
    if (lr.score < 0 || lr.n_ops == 0) return 0;
    int n_cigar = ops_to_cigar(sc, 0, lr.n_ops, cigar, kSwCigarCap - 2);
    if (n_cigar < 0) return -1;
    int x = 0, y = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const int op = cigar[k] >> 14, cl = cigar[k] & 0x3fff;
        if (op == kOpM) { x += cl; y += cl; } else if (op == kOpD) x += cl; else y += cl;
    }
    if (x < 20 || y < 20) return 0;
    // path[path_len-1] = first aligned cell, path[0] = last: walk the ops back from (end_i, end_j)
    int pi = lr.end_i, pj = lr.end_j;
    for (int k = 0; k < lr.n_ops - 1; ++k) {
        uint32_t op = sc.B(k);
        if (op == kOpM) { --pi; --pj; } else if (op == kOpI) --pj; else --pi;
    }
    *beg += (pi ? pi : 1) - 1;
    const int start = (pj ? pj : 1) - 1, end = lr.end_j;
    if (start) { for (int k = n_cigar; k > 0; --k) cigar[k] = cigar[k - 1]; cigar[0] = (uint16_t)(kOpS << 14 | start); ++n_cigar; }
    if (end < len) cigar[n_cigar++] = (uint16_t)(kOpS << 14 | (len - end));
    int n_mm = 0, n_gapo = 0, n_gape = 0;
    int px = pi ? pi - 1 : 0, py = pj ? pj - 1 : 0;
    for (int k = 0; k < n_cigar; ++k) {
        const int op = cigar[k] >> 14, cl = cigar[k] & 0x3fff;
        if (op == kOpM) {
            for (int z = 0; z < cl; ++z) { uint32_t a = R.at(px + z), b = Q.at(py + z); if (a < 4 && b < 4 && a != b) ++n_mm; }
            px += cl; py += cl;
        } else if (op == kOpD) { px += cl; ++n_gapo; n_gape += cl - 1; }
        else if (op == kOpI) { py += cl; ++n_gapo; n_gape += cl - 1; }
    }
    *cnt = (uint32_t)n_mm << 16 | (uint32_t)n_gapo << 8 | (uint32_t)n_gape;
    return n_cigar;
}

// bwa_sw_core (libbwa/bwape.c:359-445).  Returns n_cigar (0 = none), -1 = scratch / cigar capacity exceeded.
FQB_HD int sw_core(int64_t l_pac, const uint8_t *pac, const ReadSeq &Q, int64_t *beg, int reglen, uint16_t *cigar, uint32_t *cnt,
                   const DpScratch &sc) {
    const int len = Q.len;
    if (reglen < 20 || l_pac - *beg < len) return 0;
    int nn = 0;
    for (int k = 0; k < len; ++k) nn += Q.at(k) >= 4;
    if ((float)nn / len >= 0.25f || len - nn < 20) return 0;
    RefWin R; R.pac = pac; R.beg = *beg;
    { int64_t e = *beg + reglen < l_pac ? *beg + reglen : l_pac; R.l = (int)(e - *beg); }
    LocalResult lr = local_align(R, R.l, Q, len, sc, 0);
    if (lr.too_big) return -1;
    return sw_post(R, Q, lr, beg, cigar, cnt, sc);
}
struct ThreadSwCore {
    const DpScratch &sc;
    FQB_HD int operator()(int64_t l_pac, const uint8_t *pac, const ReadSeq &Q, int64_t *beg, int reglen, uint16_t *cigar, uint32_t *cnt) const {
        return sw_core(l_pac, pac, Q, beg, reglen, cigar, cnt, sc);
    }
};

// per-batch constants of bwa_paired_sw that go through libm (host, glibc): SURVEY A.5
struct SwParams {
    double avg, std;
    double s_old_add;      // -4.343 * log(ap_prior / l_pac)
    int s_new_add;         // (int)(-4.343 * log(.5 * erfc(M_SQRT1_2 * 1.5) + .499))
    int64_t l_pac;
    int on;                // popt->is_sw && ii.avg >= 0 (src/BwtMapper.cpp:1975): bwa_paired_sw runs for this batch
};

// bwa_paired_sw for one pair (libbwa/bwape.c:497-617), BWA_PET_STD.  Returns false if scratch was too small.
// __set_rght_coor / __set_left_coor of bwa_paired_sw (libbwa/bwape.c:510-523): where to look for `pm` given its mate `pref`
FQB_HD void sw_window(const fqb_read_t &pref_, const fqb_read_t &pm_, const SwParams &sp, int64_t &b, int64_t &e, int &strand) {
    const fqb_read_t *pref = &pref_, *pm = &pm_;
    const double dlen = (double)pm->len;
    if (pref->strand == 0) {        // mate on the reverse strand, to the right
        double a = FQB_DSUB(FQB_DSUB(FQB_DADD((double)(int64_t)pref->pos, sp.avg), FQB_DMUL(3.0, sp.std)), FQB_DMUL(dlen, 1.5));
        b = (int64_t)a;
        e = (int64_t)FQB_DADD(FQB_DADD((double)b, FQB_DMUL(6.0, sp.std)), (double)(2 * pm->len));
        // the reference assigns `_pref->pos + _pref->len` here, a 32-bit sum: it wraps when the gapped read's provisional
        // position underflowed (a forward read with an insertion at the very start of the first contig)
        if (b < (int64_t)pref->pos + pref->len) b = (int64_t)(uint32_t)(pref->pos + (uint32_t)pref->len);
        if (e > sp.l_pac) e = sp.l_pac;
        strand = 1;
    } else {                        // mate on the forward strand, to the left
        double a = FQB_DSUB(FQB_DSUB(FQB_DSUB((double)((int64_t)pref->pos + pref->len), sp.avg), FQB_DMUL(3.0, sp.std)), FQB_DMUL(dlen, 0.5));
        b = (int64_t)a;
        e = (int64_t)FQB_DADD(FQB_DADD((double)b, FQB_DMUL(6.0, sp.std)), (double)(2 * pm->len));
        if (b < 0) b = 0;
        if (e > (int64_t)pref->pos) e = pref->pos;
        strand = 0;
    }
}

// `core` runs bwa_sw_core for one mate (per thread, or cooperatively by a warp whose lanes all call this function).
template <class Core>
FQB_HD bool paired_sw_pair(const uint8_t *pac, fqb_read_t *p0, fqb_read_t *p1, const uint8_t *fwd0, const uint8_t *fwd1,
                           const SwParams &sp, const Core &core) {
    fqb_read_t *p[2] = {p0, p1};
    int n_cigar[2] = {0, 0}, mq_adjust[2] = {255, 255}, mapQ = 0;
    int64_t beg[2] = {0, 0};
    uint16_t cig[2][kSwCigarCap];
    uint32_t cnt[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
        const fqb_read_t *pref = p[1 - k];
        fqb_read_t *pm = p[k];
        if (pref->type == kTypeNoMatch) continue;
        ReadSeq Q; Q.fwd = k ? fwd1 : fwd0; Q.len = pm->len;
        int64_t b, e;
        sw_window(*pref, *pm, sp, b, e, Q.strand);
        beg[k] = b;
        int nc = core(sp.l_pac, pac, Q, &beg[k], (int)(e - b), cig[k], &cnt[k]);
        if (nc < 0) return false;
        n_cigar[k] = nc;
        if (nc && pm->type != kTypeNoMatch) {
            int clip = 0;
            if ((cig[k][0] >> 14) == kOpS) clip += cig[k][0] & 0x3fff;
            if ((cig[k][nc - 1] >> 14) == kOpS) clip += cig[k][nc - 1] & 0x3fff;
            int s_old = (int)(FQB_DADD(FQB_DMUL((double)(pm->n_mm * 9 + pm->n_gapo * 13 + pm->n_gape * 2) / 3., 8.), .499));
            int s_new = (int)(FQB_DADD(FQB_DMUL((double)((cnt[k] >> 16) * 9 + (cnt[k] >> 8 & 0xff) * 13 + (cnt[k] & 0xff) * 2 + (uint32_t)clip * 3) / 3., 8.), .499));
            s_old = (int)FQB_DADD((double)s_old, sp.s_old_add);
            s_new += sp.s_new_add;
            if (s_old < s_new) { mq_adjust[k] = s_new - s_old; n_cigar[k] = 0; }
            else mq_adjust[k] = s_old - s_new;
        }
    }
    int k = -1;
    if (n_cigar[0] && n_cigar[1]) { k = p0->mapQ < p1->mapQ ? 0 : 1; int d = (int)p1->mapQ - (int)p0->mapQ; mapQ = d < 0 ? -d : d; }
    else if (n_cigar[0]) { k = 0; mapQ = p1->mapQ; }
    else if (n_cigar[1]) { k = 1; mapQ = p0->mapQ; }
    if (k >= 0 && (int64_t)p[k]->pos != beg[k]) {
        fqb_read_t *pk = p[k], *po = p[1 - k];
        int tmp = (int)po->mapQ - pk->mapQ / 2 - 8;
        if (tmp <= 0) tmp = 1;
        if (mapQ > tmp) mapQ = tmp;
        pk->mapQ = po->mapQ = (uint8_t)mapQ;
        pk->seQ = po->seQ = (uint8_t)(po->seQ < mapQ ? po->seQ : mapQ);
        if (pk->mapQ > mq_adjust[k]) pk->mapQ = (uint8_t)mq_adjust[k];
        if (pk->seQ > mq_adjust[k]) pk->seQ = (uint8_t)mq_adjust[k];
        if (n_cigar[k] > FQB_MAX_CIGAR) return false;
        pk->n_cigar = (uint8_t)n_cigar[k]; pk->has_cigar = 1;
        for (int c = 0; c < n_cigar[k]; ++c) pk->cigar[c] = cig[k][c];
        pk->type = kTypeMateSW;
        pk->pos = (uint32_t)beg[k];
        pk->seQ = po->seQ;
        pk->strand = (uint8_t)(1 - po->strand);
        pk->n_mm = (uint8_t)(cnt[k] >> 16); pk->n_gapo = (uint8_t)(cnt[k] >> 8 & 0xff); pk->n_gape = (uint8_t)(cnt[k] & 0xff);
        pk->extra_flag |= kSamProper;
        po->extra_flag |= kSamProper;
    }
    return true;
}
FQB_HD bool paired_sw_one(const uint8_t *pac, fqb_read_t *p0, fqb_read_t *p1, const uint8_t *fwd0, const uint8_t *fwd1,
                          const SwParams &sp, const DpScratch &sc) {
    ThreadSwCore core{sc};
    return paired_sw_pair(pac, p0, p1, fwd0, fwd1, sp, core);
}

}  // namespace fqb
