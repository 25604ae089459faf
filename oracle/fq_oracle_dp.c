/* TEST INFRASTRUCTURE ONLY -- see fq_oracle.h.  Plain-C restatement of the banded global
 * alignment used by refine_gapped_core (libbwa/stdaln.c:345-524, aln_global_core), the CIGAR
 * construction (stdaln.c:1010-1040, bwtaln.c:352-362), refine_gapped_core itself
 * (libbwa/bwase.c:183-232), bwa_cal_md1's NM count (bwase.c:234-296) and bwa_correct_trimmed
 * (bwase.c:298-337).  Pinned against oracle/_ref. */
#include "fq_oracle.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-1073741823)
enum { T_M = 0, T_I = 1, T_D = 2 };

typedef struct { int M, I, D; } cell_t;
typedef struct { unsigned char mt, it, dt; } tr_t;

static const int kMaq[25] = {11, -19, -19, -19, -13, -19, 11, -19, -19, -13, -19, -19, 11, -19, -13,
                             -19, -19, -19, 11, -13, -13, -13, -13, -13, -13};

/* set_M / set_I / set_D and their gap_end variants (stdaln.c:260-318) */
static void from_diag(int *out, unsigned char *t, const cell_t *p, int sc)
{
    if (p->M >= p->I) { if (p->M >= p->D) { *out = p->M + sc; *t = T_M; } else { *out = p->D + sc; *t = T_D; } }
    else { if (p->I > p->D) { *out = p->I + sc; *t = T_I; } else { *out = p->D + sc; *t = T_D; } }
}
static void gap_from(int *out, unsigned char *t, int pM, int pG, int go, int ge, unsigned char self)
{
    if (pM - go > pG) { *t = T_M; *out = pM - go - ge; } else { *t = self; *out = pG - ge; }
}

/* seq1 = reference window (len1), seq2 = read (len2); ops_out receives the path's ctype sequence from the
 * END of the alignment to its start (path[0..path_len-1]); returns the score, *n_ops = path_len */
int orc_global_align(const uint8_t *seq1, int len1, const uint8_t *seq2, int len2, int gap_open, int gap_ext, int gap_end,
                     int band, uint8_t *ops_out, int *n_ops)
{
    int i, j, b1, b2, end, tmp_end, width = len1 + 1;
    cell_t *curr, *last, *s, *sw;
    tr_t *tr;
    const int *mat;
    if (len1 == 0 || len2 == 0) { *n_ops = 0; return 0; }
    if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; } else { b1 = band; b2 = len2 - len1 + band; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    --seq1; --seq2;
    tr = (tr_t *)calloc((size_t)(len2 + 1) * width, sizeof(tr_t));
    curr = (cell_t *)malloc(sizeof(cell_t) * width);
    last = (cell_t *)malloc(sizeof(cell_t) * width);
#define TR(jj, ii) (tr + (size_t)(jj) * width + (ii))
#define INF(c) ((c).M = (c).I = (c).D = NEG_INF)
#define END_GE (gap_end >= 0 ? gap_end : gap_ext)
    INF(curr[0]); curr[0].M = 0;
    for (i = 1; i < b1; ++i) { INF(curr[i]); gap_from(&curr[i].D, &TR(0, i)->dt, curr[i - 1].M, curr[i - 1].D, gap_open, END_GE, T_D); }
    sw = curr; curr = last; last = sw;
    tmp_end = (b2 < len2) ? b2 : len2 - 1;
    for (j = 1; j <= tmp_end + 1; ++j) {
        /* rows whose band starts at column 0; the extra iteration is the "last row for part 1" */
        int last_row = (j == tmp_end + 1);
        int d_end = last_row ? END_GE : gap_ext;
        if (last_row && !(j == len2 && b2 != len2 - 1)) break;
        s = curr; INF(*s);
        gap_from(&s->I, &TR(j, 0)->it, last[0].M, last[0].I, gap_open, END_GE, T_I);
        end = (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1;
        mat = kMaq + seq2[j] * 5;
        for (i = 1; i != end; ++i) {
            s = curr + i;
            from_diag(&s->M, &TR(j, i)->mt, last + i - 1, mat[seq1[i]]);
            gap_from(&s->I, &TR(j, i)->it, last[i].M, last[i].I, gap_open, gap_ext, T_I);
            gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, d_end, T_D);
        }
        s = curr + i;
        from_diag(&s->M, &TR(j, i)->mt, last + i - 1, mat[seq1[i]]);
        gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, d_end, T_D);
        if (j + b1 - 1 > len1) gap_from(&s->I, &TR(j, i)->it, last[i].M, last[i].I, gap_open, END_GE, T_I);
        else s->I = NEG_INF;
        sw = curr; curr = last; last = sw;
    }
    for (; j <= len2 - b2 + 1; ++j) {           /* band interior: both edges inside the matrix */
        INF(curr[j - b2]);
        mat = kMaq + seq2[j] * 5;
        end = j + b1 - 1;
        for (i = j - b2 + 1; i != end; ++i) {
            s = curr + i;
            from_diag(&s->M, &TR(j, i)->mt, last + i - 1, mat[seq1[i]]);
            gap_from(&s->I, &TR(j, i)->it, last[i].M, last[i].I, gap_open, gap_ext, T_I);
            gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, gap_ext, T_D);
        }
        s = curr + i;
        from_diag(&s->M, &TR(j, i)->mt, last + i - 1, mat[seq1[i]]);
        gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, gap_ext, T_D);
        s->I = NEG_INF;
        sw = curr; curr = last; last = sw;
    }
    for (; j <= len2; ++j) {                    /* band reaches the last column; j == len2 is the last row */
        int d_end = (j == len2) ? END_GE : gap_ext;
        INF(curr[j - b2]);
        mat = kMaq + seq2[j] * 5;
        for (i = j - b2 + 1; i < len1; ++i) {
            s = curr + i;
            from_diag(&s->M, &TR(j, i)->mt, last + i - 1, mat[seq1[i]]);
            gap_from(&s->I, &TR(j, i)->it, last[i].M, last[i].I, gap_open, gap_ext, T_I);
            gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, d_end, T_D);
        }
        s = curr + i;
        from_diag(&s->M, &TR(j, i)->mt, last + len1 - 1, mat[seq1[i]]);
        gap_from(&s->I, &TR(j, i)->it, last[i].M, last[i].I, gap_open, END_GE, T_I);
        gap_from(&s->D, &TR(j, i)->dt, s[-1].M, s[-1].D, gap_open, d_end, T_D);
        sw = curr; curr = last; last = sw;
    }
    {   /* backtrace */
        int max, n = 0;
        unsigned char type, ctype;
        tr_t *q;
        i = len1; j = len2;
        q = TR(j, i); s = last + len1;
        max = s->M; type = q->mt; ctype = T_M;
        if (s->I > max) { max = s->I; type = q->it; ctype = T_I; }
        if (s->D > max) { max = s->D; type = q->dt; ctype = T_D; }
        ops_out[n++] = ctype;
        do {
            if (ctype == T_M) { --i; --j; } else if (ctype == T_I) --j; else --i;
            q = TR(j, i);
            ctype = type;
            type = ctype == T_M ? q->mt : ctype == T_I ? q->it : q->dt;
            ops_out[n++] = ctype;
        } while (i || j);
        *n_ops = n - 1;
        free(tr); free(curr); free(last);
        return max;
    }
}

/* aln_path2cigar32 + bwa_aln_path2cigar: runs of ops from the start of the alignment, op<<14 | len */
int orc_ops_to_cigar(const uint8_t *ops, int n_ops, uint16_t *cigar)
{
    int i, n = 0;
    if (n_ops == 0) return 0;
    cigar[0] = (uint16_t)(ops[n_ops - 1] << 14 | 1);
    for (i = n_ops - 2; i >= 0; --i) {
        if (ops[i] == (cigar[n] >> 14)) cigar[n] += 1;
        else cigar[++n] = (uint16_t)(ops[i] << 14 | 1);
    }
    return n + 1;
}

static int pac_base(const uint8_t *pac, int64_t k) { return pac[k >> 2] >> ((~k & 3) << 1) & 3; }

/* refine_gapped_core with is_end_correct = 1 (libbwa/bwase.c:183-232) */
int orc_refine_gapped(int64_t l_pac, const uint8_t *pac, int len, const uint8_t *seq, uint32_t *pos_io, int ext, uint16_t *cigar)
{
    uint8_t ref[1024], ops[2048];
    int l = 0, n_ops, n_cigar, ref_len = len + abs(ext), kk;
    int64_t k, pos = *pos_io > (uint32_t)l_pac ? (int64_t)(int32_t)*pos_io : (int64_t)*pos_io;
    if (ext > 0) {
        for (k = pos; k < pos + ref_len && k < l_pac; ++k) ref[l++] = (uint8_t)pac_base(pac, k);
    } else {
        int64_t x = pos + len;
        for (k = x - ref_len > 0 ? x - ref_len : 0; k < x && k < l_pac; ++k) ref[l++] = (uint8_t)pac_base(pac, k);
    }
    orc_global_align(ref, l, seq, len, 26, 9, 5, 50, ops, &n_ops);
    n_cigar = orc_ops_to_cigar(ops, n_ops, cigar);
    if (ext < 0) {
        int d = 0;
        for (kk = 0; kk < n_cigar; ++kk) {
            if ((cigar[kk] >> 14) == T_D) d -= cigar[kk] & 0x3fff;
            else if ((cigar[kk] >> 14) == T_I) d += cigar[kk] & 0x3fff;
        }
        pos += d;
    }
    if ((cigar[0] >> 14) == T_D) {
        pos += cigar[0] & 0x3fff;
        for (kk = 0; kk < n_cigar - 1; ++kk) cigar[kk] = cigar[kk + 1];
        --n_cigar;
    }
    if ((cigar[n_cigar - 1] >> 14) == T_D) --n_cigar;
    if ((cigar[n_cigar - 1] >> 14) == T_I) cigar[n_cigar - 1] = (uint16_t)(3 << 14 | (cigar[n_cigar - 1] & 0x3fff));
    if ((cigar[0] >> 14) == T_I) cigar[0] = (uint16_t)(3 << 14 | (cigar[0] & 0x3fff));
    *pos_io = (uint32_t)pos;
    return n_cigar;
}

/* NM as bwa_cal_md1 counts it (libbwa/bwase.c:234-296) */
int orc_cal_nm(int n_cigar, const uint16_t *cigar, int has_cigar, int len, uint32_t pos, const uint8_t *seq, int64_t l_pac, const uint8_t *pac)
{
    int nm = 0, k, z;
    uint32_t x = pos, y = 0;
    if (has_cigar) {
        for (k = 0; k < n_cigar; ++k) {
            int l = cigar[k] & 0x3fff, op = cigar[k] >> 14;
            if (op == 0) {
                for (z = 0; z < l && (int64_t)x + z < l_pac; ++z) {
                    int c = pac_base(pac, (int64_t)x + z);
                    if (seq[y + z] > 3 || c != seq[y + z]) ++nm;
                }
                x += l; y += l;
            } else if (op == 1 || op == 3) { y += l; if (op == 1) nm += l; }
            else { x += l; nm += l; }
        }
    } else {
        for (z = 0; z < len; ++z) {
            int c = pac_base(pac, (int64_t)x + z);
            if (seq[y + z] > 3 || c != seq[y + z]) ++nm;
        }
    }
    return nm;
}

/* ---- local alignment: aln_local_core with _thres = 1, _subo = 0 (libbwa/stdaln.c:529-757).
 * Scores stay far below LOCAL_OVERFLOW_THRESHOLD (32000) for reads <= 256 bp (11 per match), so the
 * overflow-rescaling blocks of the reference can never run and are not restated.
 * Returns score_f (or -1); path as (i, j, ctype) triplets from END to START, *path_len. */
typedef struct { int i, j; unsigned char ctype; } orc_path_t;

int orc_local_align(const uint8_t *seq1_, int len1, const uint8_t *seq2_, int len2, int gap_open, int gap_ext, int band,
                    orc_path_t *path, int *path_len)
{
    const uint8_t *seq1 = seq1_ - 1, *seq2 = seq2_ - 1;
    const int q = gap_open, r = gap_ext, qr = q + r, max_score = 11;
    int *eh_h, *eh_e;          /* the packed eh[] of the reference: high half = h, low half = e */
    int i, j, score_f = 0, end_i = 0, end_j = 0, start_i = 0, start_j = 0, score_r, start, end;
    if (len1 == 0 || len2 == 0) return -1;
    eh_h = (int *)calloc((size_t)len1 + 2, sizeof(int));
    eh_e = (int *)calloc((size_t)len1 + 2, sizeof(int));
    /* forward pass */
    for (j = 1; j <= len2; ++j) {
        int last_h = 0, f = 0;
        const int *row = kMaq + seq2[j] * 5;
        for (i = 1; i <= len1; ++i) {
            int *sh = eh_h + (i - 1), *se = eh_e + (i - 1);
            int curr_h = *sh + row[seq1[i]], e;
            if (curr_h < 0) curr_h = 0;
            if (last_h > 0) { f = (f > last_h - q) ? f - r : last_h - qr; if (curr_h < f) curr_h = f; }
            if (sh[1] >= qr + 1) {                 /* *(s+1) >= (qr+1)<<16  <=>  h of the next cell >= qr+1 */
                int curr_last_h = sh[1];
                e = (*se > curr_last_h - q) ? *se - r : curr_last_h - qr;
                if (curr_h < e) curr_h = e;
                *sh = last_h; *se = e;
            } else { *sh = last_h; *se = 0; }
            last_h = curr_h;
            if (score_f < curr_h) { score_f = curr_h; end_i = i; end_j = j; }
        }
        eh_h[len1] = last_h; eh_e[len1] = 0;
    }
    if (score_f < 1) { *path_len = 0; free(eh_h); free(eh_e); return score_f; }
    /* reverse pass */
    for (i = end_i; i >= 0; --i) { eh_h[i] = 0; eh_e[i] = 0; }
    if (end_i == 0 || end_j == 0) { free(eh_h); free(eh_e); return score_f; }
    score_r = kMaq[seq1[end_i] * 5 + seq2[end_j]];
    start_i = end_i; start_j = end_j;
    eh_h[end_i] = qr + score_r; eh_e[end_i] = 0;
    start = end_i - 1;
    end = end_i - 3;
    if (end <= 0) end = 0;
    for (j = end_j - 1; j != 0; --j) {
        int last_h = 0, f = 0, stop = 0;
        const int *row = kMaq + seq2[j] * 5;
        for (i = start; i != end; --i) {
            int *sh = eh_h + (i + 1), *se = eh_e + (i + 1);
            int curr_h = *sh + row[seq1[i]], curr_last_h, e;
            if (curr_h < 0) curr_h = 0;
            if (last_h > 0) { f = (f > last_h - q) ? f - r : last_h - qr; if (curr_h < f) curr_h = f; }
            curr_last_h = sh[-1];
            e = (*se > curr_last_h - q) ? *se - r : curr_last_h - qr;
            if (e < 0) e = 0;
            if (curr_h < e) curr_h = e;
            *sh = last_h; *se = e;
            last_h = curr_h;
            if (score_r < curr_h) {
                score_r = curr_h; start_i = i; start_j = j;
                if (score_r - qr == score_f) { stop = 1; break; }
            }
        }
        if (stop) break;           /* the reference sets j = 1 and breaks: the for's --j then ends the loop */
        eh_h[i + 1] = last_h; eh_e[i + 1] = 0;
        if (eh_h[start] <= qr) --start;
        if (start <= 0) start = 0;
        end = start_i - (start_j - j) - (score_r + (start_j - j) * max_score) / r - 1;
        if (end <= 0) end = 0;
    }
    score_r -= qr;
    {   /* global alignment inside [start, end] with gap_end = -1, widening the band until the scores agree */
        int score_g, w, n_ops = 0, k, jmax = (end_i - start_i > end_j - start_j) ? end_i - start_i : end_j - start_j;
        uint8_t *ops = (uint8_t *)malloc((size_t)len1 + len2 + 8);
        ++jmax;
        for (w = band;; w <<= 1) {
            score_g = orc_global_align(seq1 + start_i, end_i - start_i + 1, seq2 + start_j, end_j - start_j + 1, gap_open, gap_ext, -1, w, ops, &n_ops);
            if (score_g == score_r || score_f == score_g) break;
            if (w > jmax) break;
        }
        if (score_r > score_g && score_f > score_g) score_f = -1;
        else score_f = score_g;
        /* rebuild (i, j) along the path, then shift into window coordinates */
        {
            int ci = end_i - start_i + 1, cj = end_j - start_j + 1;
            for (k = 0; k < n_ops; ++k) {
                path[k].i = ci + start_i - 1; path[k].j = cj + start_j - 1; path[k].ctype = ops[k];
                if (ops[k] == T_M) { --ci; --cj; } else if (ops[k] == T_I) --cj; else --ci;
            }
        }
        *path_len = n_ops;
        free(ops);
    }
    free(eh_h); free(eh_e);
    return score_f;
}

/* bwa_sw_core (libbwa/bwape.c:359-445).  Returns n_cigar (0 = no alignment); *cnt = n_mm<<16 | n_gapo<<8 | n_gape */
int orc_sw_core(int64_t l_pac, const uint8_t *pac, int len, const uint8_t *seq, int64_t *beg, int reglen, uint16_t *cigar, uint32_t *cnt)
{
    uint8_t *ref;
    orc_path_t *path;
    uint8_t *ops;
    int64_t k;
    int l = 0, x = 0, y, path_len = 0, ret, n_cigar, kk, start, end;
    if (reglen < 20 || l_pac - *beg < len) return 0;
    for (kk = 0; kk < len; ++kk) if (seq[kk] >= 4) ++x;
    if ((float)x / len >= 0.25 || len - x < 20) return 0;
    ref = (uint8_t *)calloc((size_t)reglen + 1, 1);
    for (k = *beg; l < reglen && k < l_pac; ++k) ref[l++] = (uint8_t)pac_base(pac, k);
    path = (orc_path_t *)calloc((size_t)l + len + 8, sizeof(orc_path_t));
    ret = orc_local_align(ref, l, seq, len, 26, 9, 50, path, &path_len);
    if (ret < 0 || path_len == 0) { free(path); free(ref); return 0; }
    ops = (uint8_t *)malloc((size_t)path_len);
    for (kk = 0; kk < path_len; ++kk) ops[kk] = path[kk].ctype;
    n_cigar = orc_ops_to_cigar(ops, path_len, cigar);
    free(ops);
    for (kk = 0, x = y = 0; kk < n_cigar; ++kk) {
        int op = cigar[kk] >> 14, cl = cigar[kk] & 0x3fff;
        if (op == T_M) { x += cl; y += cl; } else if (op == T_D) x += cl; else y += cl;
    }
    if (x < 20 || y < 20) { free(path); free(ref); return 0; }
    {
        const orc_path_t *p = path + path_len - 1;
        int px = p->i ? p->i - 1 : 0, py = p->j ? p->j - 1 : 0, n_mm = 0, n_gapo = 0, n_gape = 0, z;
        *beg += (p->i ? p->i : 1) - 1;
        start = (p->j ? p->j : 1) - 1;
        end = path->j;
        if (start) { memmove(cigar + 1, cigar, sizeof(uint16_t) * (size_t)n_cigar); cigar[0] = (uint16_t)(3 << 14 | start); ++n_cigar; }
        if (end < len) cigar[n_cigar++] = (uint16_t)(3 << 14 | (len - end));
        for (kk = 0; kk < n_cigar; ++kk) {
            int op = cigar[kk] >> 14, cl = cigar[kk] & 0x3fff;
            if (op == T_M) {
                for (z = 0; z < cl; ++z) if (ref[px + z] < 4 && seq[py + z] < 4 && ref[px + z] != seq[py + z]) ++n_mm;
                px += cl; py += cl;
            } else if (op == T_D) { px += cl; ++n_gapo; n_gape += cl - 1; }
            else if (op == T_I) { py += cl; ++n_gapo; n_gape += cl - 1; }
        }
        *cnt = (uint32_t)n_mm << 16 | (uint32_t)n_gapo << 8 | (uint32_t)n_gape;
    }
    free(path); free(ref);
    return n_cigar;
}

/* ---- batch drivers ------------------------------------------------------------------------- */
#include <math.h>
#define TYPE_NO_MATCH 0
#define TYPE_MATESW 3
#define F_PROPER 2

/* read r in alignment orientation: strand 0 -> the read as sequenced, strand 1 -> its reverse complement */
static void oriented(const uint8_t *fwd, int len, int strand, uint8_t *out)
{
    int j;
    for (j = 0; j < len; ++j) {
        uint8_t c = strand ? fwd[len - 1 - j] : fwd[j];
        out[j] = (strand && c < 4) ? (uint8_t)(3 - c) : c;
    }
}

/* bwa_paired_sw (libbwa/bwape.c:463-625), BWA_PET_STD only */
void orc_paired_sw(int64_t l_pac, const uint8_t *pac, int n_pairs, orc_row_t *rows, const uint8_t *codes, int stride,
                   const orc_pe_opt_t *popt, const orc_isize_t *ii)
{
    int i, k;
    if (!popt->is_sw || ii->avg < 0.0) return;
    for (i = 0; i < n_pairs; ++i) {
        orc_row_t *p[2] = {rows + 2 * i, rows + 2 * i + 1};
        if (p[0]->filtered) { if (p[1]->filtered) continue; p[0]->filtered = 0; }
        else if (p[1]->filtered) p[1]->filtered = 0;
        if ((p[0]->mapQ >= 17 || p[1]->mapQ >= 17) && (p[0]->extra_flag & F_PROPER) == 0) {
            int n_cigar[2] = {0, 0}, mapQ = 0, mq_adjust[2] = {255, 255};
            int64_t beg[2] = {0, 0}, end[2] = {0, 0};
            uint16_t cigar[2][64];
            uint32_t cnt[2] = {0, 0};
            for (k = 0; k < 2; ++k) {
                const orc_row_t *pref = p[1 - k];
                orc_row_t *pm = p[k];
                uint8_t seq[1024];
                if (pref->type == TYPE_NO_MATCH) continue;
                if (pref->strand == 0) {
                    beg[k] = (int64_t)((int64_t)pref->pos + ii->avg - 3 * ii->std - pm->len * 1.5);
                    end[k] = (int64_t)(beg[k] + 6 * ii->std + 2 * pm->len);
                    if (beg[k] < (int64_t)pref->pos + pref->len) beg[k] = (int64_t)(uint32_t)(pref->pos + (uint32_t)pref->len);   /* 32-bit sum in the reference (bwape.c:513) */
                    if (end[k] > l_pac) end[k] = l_pac;
                    oriented(codes + (size_t)(2 * i + k) * stride, pm->len, 1, seq);
                } else {
                    beg[k] = (int64_t)((int64_t)pref->pos + pref->len - ii->avg - 3 * ii->std - pm->len * 0.5);
                    end[k] = (int64_t)(beg[k] + 6 * ii->std + 2 * pm->len);
                    if (beg[k] < 0) beg[k] = 0;
                    if (end[k] > (int64_t)pref->pos) end[k] = pref->pos;
                    oriented(codes + (size_t)(2 * i + k) * stride, pm->len, 0, seq);
                }
                n_cigar[k] = orc_sw_core(l_pac, pac, pm->len, seq, &beg[k], (int)(end[k] - beg[k]), cigar[k], &cnt[k]);
                if (n_cigar[k] && pm->type != TYPE_NO_MATCH) {
                    int s_old, s_new, clip = 0;
                    if ((cigar[k][0] >> 14) == 3) clip += cigar[k][0] & 0x3fff;
                    if ((cigar[k][n_cigar[k] - 1] >> 14) == 3) clip += cigar[k][n_cigar[k] - 1] & 0x3fff;
                    s_old = (int)((pm->n_mm * 9 + pm->n_gapo * 13 + pm->n_gape * 2) / 3. * 8. + .499);
                    s_new = (int)(((cnt[k] >> 16) * 9 + (cnt[k] >> 8 & 0xff) * 13 + (cnt[k] & 0xff) * 2 + clip * 3) / 3. * 8. + .499);
                    s_old += -4.343 * log(ii->ap_prior / l_pac);
                    s_new += (int)(-4.343 * log(.5 * erfc(M_SQRT1_2 * 1.5) + .499));
                    if (s_old < s_new) { mq_adjust[k] = s_new - s_old; n_cigar[k] = 0; }
                    else mq_adjust[k] = s_old - s_new;
                }
            }
            k = -1;
            if (n_cigar[0] && n_cigar[1]) { k = p[0]->mapQ < p[1]->mapQ ? 0 : 1; mapQ = abs((int)p[1]->mapQ - (int)p[0]->mapQ); }
            else if (n_cigar[0]) { k = 0; mapQ = p[1]->mapQ; }
            else if (n_cigar[1]) { k = 1; mapQ = p[0]->mapQ; }
            if (k >= 0 && (int64_t)p[k]->pos != beg[k]) {
                int tmp = (int)p[1 - k]->mapQ - p[k]->mapQ / 2 - 8, c;
                if (tmp <= 0) tmp = 1;
                if (mapQ > tmp) mapQ = tmp;
                p[k]->mapQ = p[1 - k]->mapQ = (uint8_t)mapQ;
                p[k]->seQ = p[1 - k]->seQ = (uint8_t)(p[1 - k]->seQ < mapQ ? p[1 - k]->seQ : mapQ);
                if (p[k]->mapQ > mq_adjust[k]) p[k]->mapQ = (uint8_t)mq_adjust[k];
                if (p[k]->seQ > mq_adjust[k]) p[k]->seQ = (uint8_t)mq_adjust[k];
                p[k]->n_cigar = (uint8_t)n_cigar[k]; p[k]->has_cigar = 1;
                for (c = 0; c < n_cigar[k] && c < 24; ++c) p[k]->cigar[c] = cigar[k][c];
                p[k]->type = TYPE_MATESW;
                p[k]->pos = (uint32_t)beg[k];
                p[k]->seQ = p[1 - k]->seQ;
                p[k]->strand = (uint8_t)(1 - p[1 - k]->strand);
                p[k]->n_mm = (uint8_t)(cnt[k] >> 16); p[k]->n_gapo = (uint8_t)(cnt[k] >> 8 & 0xff); p[k]->n_gape = (uint8_t)(cnt[k] & 0xff);
                p[k]->extra_flag |= F_PROPER;
                p[1 - k]->extra_flag |= F_PROPER;
            }
        }
    }
}

/* bwa_refine_gapped for one end's reads (libbwa/bwase.c:339-418): CIGAR + position fix-up, NM, trim correction */
void orc_refine_gapped_batch(int64_t l_pac, const uint8_t *pac, int n_reads, orc_row_t *rows, const uint8_t *codes, int stride)
{
    int r;
    for (r = 0; r < n_reads; ++r) {
        orc_row_t *s = rows + r;
        uint8_t seq[1024];
        if (!s->filtered && !(s->type == TYPE_NO_MATCH || s->type == TYPE_MATESW || s->n_gapo == 0)) {
            oriented(codes + (size_t)r * stride, s->len, s->strand, seq);
            s->n_cigar = (uint8_t)orc_refine_gapped(l_pac, pac, s->len, seq, &s->pos, (s->strand ? 1 : -1) * (s->n_gapo + s->n_gape), s->cigar);
            s->has_cigar = 1;
        }
    }
    for (r = 0; r < n_reads; ++r) {
        orc_row_t *s = rows + r;
        uint8_t seq[1024];
        if (s->type != TYPE_NO_MATCH) {
            oriented(codes + (size_t)r * stride, s->len, s->strand, seq);
            s->nm = (uint16_t)orc_cal_nm(s->n_cigar, s->cigar, s->has_cigar, s->len, s->pos, seq, l_pac, pac);
        }
    }
    for (r = 0; r < n_reads; ++r) {      /* bwa_correct_trimmed, every read */
        orc_row_t *s = rows + r;
        int clip = s->full_len - s->len, k;
        if (clip == 0) continue;
        if (s->strand == 0) {
            if (s->has_cigar && (s->cigar[s->n_cigar - 1] >> 14) == 3) s->cigar[s->n_cigar - 1] += (uint16_t)clip;
            else {
                if (!s->has_cigar) { s->n_cigar = 2; s->has_cigar = 1; s->cigar[0] = (uint16_t)(0 << 14 | s->len); }
                else ++s->n_cigar;
                s->cigar[s->n_cigar - 1] = (uint16_t)(3 << 14 | clip);
            }
        } else {
            if (s->has_cigar && (s->cigar[0] >> 14) == 3) s->cigar[0] += (uint16_t)clip;
            else {
                if (!s->has_cigar) { s->n_cigar = 2; s->has_cigar = 1; s->cigar[1] = (uint16_t)(0 << 14 | s->len); }
                else { ++s->n_cigar; for (k = s->n_cigar - 1; k > 0; --k) s->cigar[k] = s->cigar[k - 1]; }
                s->cigar[0] = (uint16_t)(3 << 14 | clip);
            }
        }
        s->len = s->full_len;
    }
}

/* ---- MD tag and its inverse (test infrastructure for rows a11/a13) ---------------------------------------
 * orc_cal_md: the MD string and NM of bwa_cal_md1 (libbwa/bwase.c:234-296). `seq` is the read in alignment
 * orientation.  Returns NM, writes a NUL-terminated string (truncated at cap-1). */
static int md_put(char *out, int cap, int o, const char *s) { while (*s && o < cap - 1) out[o++] = *s++; out[o] = 0; return o; }
static int md_num(char *out, int cap, int o, int u) { char b[16]; snprintf(b, sizeof b, "%d", u); return md_put(out, cap, o, b); }
static int md_chr(char *out, int cap, int o, char c) { char b[2] = {c, 0}; return md_put(out, cap, o, b); }

int orc_cal_md(int n_cigar, const uint16_t *cigar, int has_cigar, int len, uint32_t pos, const uint8_t *seq, int64_t l_pac,
               const uint8_t *pac, char *out, int cap)
{
    int nm = 0, u = 0, o = 0, k, z;
    uint32_t x = pos, y = 0;
    out[0] = 0;
    if (has_cigar) {
        for (k = 0; k < n_cigar; ++k) {
            int l = cigar[k] & 0x3fff, op = cigar[k] >> 14;
            if (op == 0) {
                for (z = 0; z < l && (int64_t)x + z < l_pac; ++z) {
                    int c = pac_base(pac, (int64_t)x + z);
                    if (seq[y + z] > 3 || c != seq[y + z]) { o = md_num(out, cap, o, u); o = md_chr(out, cap, o, "ACGTN"[c]); ++nm; u = 0; }
                    else ++u;
                }
                x += l; y += l;
            } else if (op == 1 || op == 3) { y += l; if (op == 1) nm += l; }
            else {
                o = md_num(out, cap, o, u); o = md_chr(out, cap, o, '^');
                for (z = 0; z < l && (int64_t)x + z < l_pac; ++z) o = md_chr(out, cap, o, "ACGT"[pac_base(pac, (int64_t)x + z)]);
                u = 0; x += l; nm += l;
            }
        }
    } else {
        for (z = 0; z < len; ++z) {
            int c = pac_base(pac, (int64_t)x + z);
            if (seq[y + z] > 3 || c != seq[y + z]) { o = md_num(out, cap, o, u); o = md_chr(out, cap, o, "ACGTN"[c]); ++nm; u = 0; }
            else ++u;
        }
    }
    md_num(out, cap, o, u);
    return nm;
}

/* StatCollector::RecoverRefseqByMDandCigar (src/StatCollector.cpp:101-172): the reference bases under a read,
 * rebuilt from the read, its MD tag and CIGAR; quirks kept (upper-casing, the perfect-match shortcut, M pieces
 * only, deletions spliced in at the running MD coordinate).  Returns the length written (NUL-terminated). */
int orc_recover_refseq(const char *read, const char *md_in, const uint16_t *cigar, int n_cigar, char *out, int cap)
{
    char md[2048];
    int n = 0, i, k, rl = (int)strlen(read), ml;
    int last = 0, total = 0, has_base = 0;
    for (i = 0; md_in[i] && i < (int)sizeof md - 1; ++i) md[i] = (char)toupper((unsigned char)md_in[i]);
    md[i] = 0; ml = i;
    for (i = 0; i < ml; ++i) if (strchr("ATCGN", md[i])) has_base = 1;
    if (!has_base && atol(md) == (long)rl) { n = rl < cap - 1 ? rl : cap - 1; memcpy(out, read, (size_t)n); out[n] = 0; return n; }
    if (n_cigar > 0) {
        int at = 0;
        for (k = 0; k < n_cigar; ++k) {
            int cl = cigar[k] & 0x3fff, op = cigar[k] >> 14;
            if (op == 0) {                                /* substr(at, cl): clipped at the end of the read */
                int take = at >= rl ? 0 : (at + cl > rl ? rl - at : cl);
                if (n + take > cap - 1) take = cap - 1 - n;
                memcpy(out + n, read + at, (size_t)take); n += take; at += cl;
            } else if (op == 3 || op == 1) at += cl;
        }
    } else { n = rl < cap - 1 ? rl : cap - 1; memcpy(out, read, (size_t)n); }
    out[n] = 0;
    for (i = 0; i < ml; ++i) {
        if (isdigit((unsigned char)md[i])) continue;
        if (md[i] == '^') {
            char num[32], del[512];
            int len, start, dl = 0, m = i - last < 31 ? i - last : 31;
            memcpy(num, md + last, (size_t)m); num[m] = 0;
            len = atoi(num); total += len; start = total;
            ++i;
            while (md[i] && !isdigit((unsigned char)md[i]) && dl < (int)sizeof del - 1) { del[dl++] = md[i]; ++i; ++total; }
            if (start > n) start = n;
            if (n + dl > cap - 1) dl = cap - 1 - n;
            memmove(out + start + dl, out + start, (size_t)(n - start));
            memcpy(out + start, del, (size_t)dl);
            n += dl; out[n] = 0;
            last = i;
        } else {
            char num[32];
            int len, m = i - last < 31 ? i - last : 31;
            memcpy(num, md + last, (size_t)m); num[m] = 0;
            len = atoi(num) + 1; total += len;
            if (total - 1 < n) out[total - 1] = md[i];
            last = i + 1;
        }
    }
    return n;
}
