/* TEST INFRASTRUCTURE ONLY -- see fq_oracle.h.  Plain-C restatement of the paired-end
 * resolution stage: SE hit choice, insert-size inference, pairing.  Parity pinned
 * against oracle/_ref (snapshot after bwa_cal_pac_pos_pe). */
#include "fq_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TYPE_NO_MATCH 0
#define TYPE_UNIQUE 1
#define TYPE_REPEAT 2
#define F_PAIRED 1
#define F_PROPER 2
#define F_READ1 64
#define F_READ2 128

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* infer_isize, libbwa/bwape.c:49-117 */
int orc_infer_isize(int n_pairs, const orc_row_t *rows, orc_isize_t *ii, double ap_prior, int64_t L)
{
    uint64_t x, *isz, n_ap = 0;
    int n, i, tot = 0, p25, p50, p75, max_len = 1, tmp;
    double skew = 0.0, kurt = 0.0, y;
    (void)p50;
    ii->avg = ii->std = -1.0;
    ii->low = ii->high = ii->high_bayesian = 0;
    isz = (uint64_t *)calloc((size_t)n_pairs + 1, 8);
    for (i = 0; i < n_pairs; ++i) {
        const orc_row_t *p0 = rows + 2 * i, *p1 = p0 + 1;
        if (p0->mapQ >= 20 && p1->mapQ >= 20) {
            x = (p0->pos < p1->pos) ? (uint64_t)p1->pos + p1->len - p0->pos : (uint64_t)p0->pos + p0->len - p1->pos;
            if (x < 100000) isz[tot++] = x;
        }
        if (p0->len > max_len) max_len = p0->len;
        if (p1->len > max_len) max_len = p1->len;
    }
    if (tot < 20) { free(isz); return -1; }
    qsort(isz, (size_t)tot, 8, cmp_u64);
    p25 = (int)isz[(int)(tot * 0.25 + 0.5)];
    p50 = (int)isz[(int)(tot * 0.50 + 0.5)];
    p75 = (int)isz[(int)(tot * 0.75 + 0.5)];
    tmp = (int)(p25 - 2.0 * (p75 - p25) + .499);
    ii->low = tmp > max_len ? (uint32_t)tmp : (uint32_t)max_len;
    ii->high = (uint32_t)(int)(p75 + 2.0 * (p75 - p25) + .499);
    for (i = 0, x = 0, n = 0; i < tot; ++i)
        if (isz[i] >= ii->low && isz[i] <= ii->high) { ++n; x += isz[i]; }
    ii->avg = (double)x / n;
    for (i = 0; i < tot; ++i)
        if (isz[i] >= ii->low && isz[i] <= ii->high) {
            double t = (isz[i] - ii->avg) * (isz[i] - ii->avg);
            ii->std += t; skew += t * (isz[i] - ii->avg); kurt += t * t;
        }
    kurt = kurt / n / (ii->std / n * ii->std / n) - 3;
    ii->std = sqrt(ii->std / n);
    skew = skew / n / (ii->std * ii->std * ii->std);
    (void)kurt; (void)skew;
    for (y = 1.0; y < 10.0; y += 0.01)
        if (.5 * erfc(y / M_SQRT2) < ap_prior / L * (y * ii->std + ii->avg)) break;
    ii->high_bayesian = (uint32_t)(y * ii->std + ii->avg + .499);
    for (i = 0; i < tot; ++i) if (isz[i] > ii->high_bayesian) ++n_ap;
    ii->ap_prior = .01 * (n_ap + .01) / tot;
    if (ii->ap_prior < ap_prior) ii->ap_prior = ap_prior;
    free(isz);
    if (isnan(ii->std) || p75 > 100000) {
        ii->low = ii->high = ii->high_bayesian = 0; ii->avg = ii->std = -1.0;
        return -1;
    }
    for (y = 1.0; y < 10.0; y += 0.01)
        if (.5 * erfc(y / M_SQRT2) < ap_prior / L * (y * ii->std + ii->avg)) break;
    ii->high_bayesian = (uint32_t)(y * ii->std + ii->avg + .499);
    return 0;
}

/* hash_64, libbwa/bwape.h:41-52 */
static uint64_t mix64(uint64_t key)
{
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

static uint32_t hit_pos(const orc_bwt_t *const bwts[2], int a, uint32_t row, int len)
{
    return a ? orc_sa(bwts[0], row) : bwts[1]->seq_len - (orc_sa(bwts[1], row) + (uint32_t)len);
}

/* pairing + __pairing_aux + __pairing_aux2, libbwa/bwape.c:119-213, bwape.h:55-82 */
static void pair_up(orc_row_t *p[2], const orc_aln_t *aln[2], uint64_t *arr, int n_arr, const orc_pe_opt_t *opt, int s_mm,
                    const orc_isize_t *ii, const int g_log_n[256])
{
    int i, j, o_n = 0, subo_n = 0, max_len;
    uint64_t last_pos[2][2], o_pos[2] = {0, 0}, subo_score, o_score;
    max_len = p[0]->full_len;
    if (max_len < p[1]->full_len) max_len = p[1]->full_len;
    o_score = subo_score = (uint64_t)-1;
    qsort(arr, (size_t)n_arr, 8, cmp_u64);
    for (j = 0; j < 2; ++j) last_pos[j][0] = last_pos[j][1] = (uint64_t)-1;
    for (i = 0; i < n_arr; ++i) {
        uint64_t x = arr[i];
        int strand = aln[x & 1][(uint32_t)x >> 1].a;
        if (strand == 1) {
            int y = 1 - (int)(x & 1), t;
            for (t = 1; t >= 0; --t) {               /* last_pos[y][1] first, then [y][0] */
                uint64_t u = last_pos[y][t], v = x;
                uint32_t l = (uint32_t)((v >> 32) + (uint64_t)p[v & 1]->len - (u >> 32));
                if (u != (uint64_t)-1 && (v >> 32) > (u >> 32) && l >= (uint32_t)max_len
                    && ((ii->high && l <= ii->high_bayesian) || (ii->high == 0 && l <= (uint32_t)opt->max_isize))) {
                    uint64_t s = (uint64_t)(aln[v & 1][(uint32_t)v >> 1].score + aln[u & 1][(uint32_t)u >> 1].score);
                    s *= 10;
                    if (ii->high) s += (uint64_t)(int)(-4.343 * log(.5 * erfc(M_SQRT1_2 * fabs(l - ii->avg) / ii->std)) + .499);
                    s = s << 32 | (uint32_t)mix64((u >> 32 << 32) | (v >> 32));
                    if (s >> 32 == o_score >> 32) ++o_n;
                    else if (s >> 32 < o_score << 32) { subo_n += o_n; o_n = 1; }      /* sic: the shift goes the wrong way */
                    else ++subo_n;
                    if (s < o_score) { subo_score = o_score; o_score = s; o_pos[u & 1] = u; o_pos[v & 1] = v; }
                    else if (s < subo_score) subo_score = s;
                }
            }
        } else {
            last_pos[x & 1][0] = last_pos[x & 1][1];
            last_pos[x & 1][1] = x;
        }
    }
    if (o_score != (uint64_t)-1) {
        int mapQ_p = 0, rr[2];
        if (o_n == 1) {
            if (subo_score == (uint64_t)-1) mapQ_p = 29;
            else if ((subo_score >> 32) - (o_score >> 32) > (uint64_t)(s_mm * 10)) mapQ_p = 23;
            else {
                int n = subo_n > 255 ? 255 : subo_n;
                mapQ_p = (int)(((subo_score >> 32) - (o_score >> 32)) / 2) - g_log_n[n];
                if (mapQ_p < 0) mapQ_p = 0;
            }
        }
        rr[0] = aln[o_pos[0] & 1][(uint32_t)o_pos[0] >> 1].a;
        rr[1] = aln[o_pos[1] & 1][(uint32_t)o_pos[1] >> 1].a;
        if ((p[0]->pos == o_pos[0] >> 32 && p[0]->strand == rr[0]) && (p[1]->pos == o_pos[1] >> 32 && p[1]->strand == rr[1])) {
            if (p[0]->mapQ > 0 && p[1]->mapQ > 0) {
                int mapQ = p[0]->mapQ + p[1]->mapQ;
                if (mapQ > 60) mapQ = 60;
                p[0]->mapQ = p[1]->mapQ = (uint8_t)mapQ;
            } else {
                if (p[0]->mapQ == 0) p[0]->mapQ = (uint8_t)((mapQ_p + 7 < p[1]->mapQ) ? mapQ_p + 7 : p[1]->mapQ);
                if (p[1]->mapQ == 0) p[1]->mapQ = (uint8_t)((mapQ_p + 7 < p[0]->mapQ) ? mapQ_p + 7 : p[0]->mapQ);
            }
        } else if (p[0]->pos == o_pos[0] >> 32 && p[0]->strand == rr[0]) {
            p[1]->seQ = 0; p[1]->mapQ = p[0]->mapQ;
            if (p[1]->mapQ > mapQ_p) p[1]->mapQ = (uint8_t)mapQ_p;
        } else if (p[1]->pos == o_pos[1] >> 32 && p[1]->strand == rr[1]) {
            p[0]->seQ = 0; p[0]->mapQ = p[1]->mapQ;
            if (p[0]->mapQ > mapQ_p) p[0]->mapQ = (uint8_t)mapQ_p;
        } else {
            p[0]->seQ = p[1]->seQ = 0;
            mapQ_p -= 20;
            if (mapQ_p < 0) mapQ_p = 0;
            p[0]->mapQ = p[1]->mapQ = (uint8_t)mapQ_p;
        }
        for (j = 0; j < 2; ++j) {
            uint64_t w = o_pos[j];
            const orc_aln_t *r = aln[w & 1] + ((uint32_t)w >> 1);
            orc_row_t *q = p[j];
            q->extra_flag |= F_PROPER;
            if (q->pos != w >> 32 || q->strand != r->a) {
                q->n_mm = r->n_mm; q->n_gapo = r->n_gapo; q->n_gape = r->n_gape; q->strand = r->a;
                q->score = r->score; q->pos = (uint32_t)(w >> 32);
            }
        }
    }
}

/* bwa_aln2seq_core with set_main = 0 (libbwa/bwase.c:47-95): the multi-hit list in SA coordinates */
static int multi_hits(int n_aln, const orc_aln_t *aln, uint32_t main_sa, int n_multi, uint32_t *rows_out, uint8_t *strand_out)
{
    int k, z = 0, zz = 0;
    uint32_t n_occ = 0, l;
    for (k = 0; k < n_aln; ++k) n_occ += aln[k].l - aln[k].k + 1;
    if (n_occ > (uint32_t)n_multi + 1) return 0;
    for (k = 0; k < n_aln; ++k)
        for (l = aln[k].k; l <= aln[k].l; ++l) {
            if (l != main_sa) { if (zz < 16) { rows_out[zz] = l; strand_out[zz] = aln[k].a; } ++zz; }
            ++z;
        }
    (void)z;
    return zz < n_multi ? zz : n_multi;
}

int orc_cal_pac_pos_pe(const orc_bwt_t *const bwts[2], int n_pairs, orc_row_t *rows, const int32_t *n_aln,
                       const orc_aln_t *aln, int aln_cap, double fnr, int opt_max_diff, int s_mm,
                       const orc_pe_opt_t *popt, orc_rng_t *rng, const orc_isize_t *last_ii, orc_isize_t *ii,
                       uint32_t *multi_pos)
{
    int i, j, g_log_n[256];
    uint64_t *arr = 0;
    size_t arr_cap = 0;
    orc_fill_log_n(g_log_n);
    /* SE pass, src/BwtMapper.cpp:744-776 */
    for (i = 0; i < n_pairs; ++i)
        for (j = 0; j < 2; ++j) {
            int r = 2 * i + j;
            orc_row_t *p = rows + r;
            orc_se_t se;
            p->n_multi = 0;
            p->extra_flag |= F_PAIRED | (j == 0 ? F_READ1 : F_READ2);
            if (p->filtered) continue;
            memset(&se, 0, sizeof se);
            se.sa = p->sa; se.n_mm = p->n_mm; se.n_gapo = p->n_gapo; se.n_gape = p->n_gape; se.strand = p->strand; se.score = p->score;
            orc_aln2seq_main(n_aln[r], aln + (size_t)r * aln_cap, rng, &se);
            p->type = se.type; p->c1 = se.c1; p->c2 = se.c2;
            if (n_aln[r]) { p->sa = se.sa; p->n_mm = se.n_mm; p->n_gapo = se.n_gapo; p->n_gape = se.n_gape; p->strand = se.strand; p->score = se.score; }
            if (p->type == TYPE_UNIQUE || p->type == TYPE_REPEAT) {
                int max_diff = fnr > 0.0 ? orc_cal_maxdiff(p->len, 0.02, fnr) : opt_max_diff;
                p->pos = hit_pos(bwts, p->strand, p->sa, p->len);
                p->seQ = p->mapQ = (uint8_t)orc_approx_mapq(&se, max_diff, g_log_n);
            }
        }
    /* insert size, :779-786 */
    orc_infer_isize(n_pairs, rows, ii, popt->ap_prior, bwts[0]->seq_len);
    if (ii->avg < 0.0 && last_ii->avg > 0.0) *ii = *last_ii;
    if (popt->force_isize) { ii->low = ii->high = 0; ii->avg = ii->std = -1.0; }
    /* PE pass, :789-886 */
    for (i = 0; i < n_pairs; ++i) {
        orc_row_t *p[2] = {rows + 2 * i, rows + 2 * i + 1};
        const orc_aln_t *al[2] = {aln + (size_t)(2 * i) * aln_cap, aln + (size_t)(2 * i + 1) * aln_cap};
        int na[2] = {p[0]->filtered ? 0 : n_aln[2 * i], p[1]->filtered ? 0 : n_aln[2 * i + 1]};
        if ((p[0]->type == TYPE_UNIQUE || p[0]->type == TYPE_REPEAT) && (p[1]->type == TYPE_UNIQUE || p[1]->type == TYPE_REPEAT)) {
            uint32_t n_occ[2] = {0, 0}, k, l;
            for (j = 0; j < 2; ++j) for (k = 0; k < (uint32_t)na[j]; ++k) n_occ[j] += al[j][k].l - al[j][k].k + 1;
            if (!(n_occ[0] > popt->max_occ || n_occ[1] > popt->max_occ)) {
                size_t n_arr = 0;
                if (arr_cap < (size_t)n_occ[0] + n_occ[1]) { arr_cap = (size_t)n_occ[0] + n_occ[1] + 16; arr = (uint64_t *)realloc(arr, arr_cap * 8); }
                for (j = 0; j < 2; ++j)
                    for (k = 0; k < (uint32_t)na[j]; ++k)
                        for (l = al[j][k].k; l <= al[j][k].l; ++l)
                            arr[n_arr++] = (uint64_t)hit_pos(bwts, al[j][k].a, l, p[j]->len) << 32 | k << 1 | (uint32_t)j;
                pair_up(p, al, arr, (int)n_arr, popt, s_mm, ii, g_log_n);
            }
        }
        if (popt->N_multi || popt->n_multi)
            for (j = 0; j < 2; ++j)
                if (p[j]->type != TYPE_NO_MATCH) {
                    uint32_t mrow[16]; uint8_t mstr[16];
                    int nm, k, lim;
                    if (!(p[j]->extra_flag & F_PROPER) && p[1 - j]->type != TYPE_NO_MATCH)
                        lim = (p[j]->c1 + p[j]->c2 - 1 > (uint32_t)popt->N_multi) ? popt->n_multi : popt->N_multi;
                    else lim = popt->n_multi;
                    nm = multi_hits(na[j], al[j], p[j]->sa, lim, mrow, mstr);
                    p[j]->n_multi = (uint8_t)nm;
                    if (multi_pos)
                        for (k = 0; k < nm && k < 11; ++k)
                            multi_pos[(size_t)(2 * i + j) * 11 + k] = hit_pos(bwts, mstr[k], mrow[k], p[j]->len);
                }
    }
    free(arr);
    return 0;
}
