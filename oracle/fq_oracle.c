/* TEST INFRASTRUCTURE ONLY -- see fq_oracle.h.  Plain-C restatement of the
 * reference's FM-index search path.  Parity pinned against oracle/_ref. */
#include "fq_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

uint64_t orc_blk_touches = 0;
uint64_t orc_pops = 0, orc_peak_entries = 0;

#define BLK_BASES 128u
#define BLK_WORDS 12u

/* ---- rank queries (libbwa/bwt.h:89-222) ---------------------------------- */

/* number of symbols == c among the first `upto` bases (0..128) of block words */
static uint32_t count_in_block(const uint32_t *bases, uint32_t upto, int c)
{
    uint32_t n = 0, j;
    for (j = 0; j < upto; ++j)
        n += ((bases[j >> 4] >> ((15u - (j & 15u)) << 1)) & 3u) == (uint32_t)c;
    return n;
}

/* bwt_occ, libbwa/bwt.h:98-122 */
uint32_t orc_occ(const orc_bwt_t *b, uint32_t k, int c)
{
    const uint32_t *blk;
    ++orc_blk_touches;
    if (k == b->seq_len) return b->L2[c + 1] - b->L2[c];
    if (k == 0xffffffffu) return 0;
    if (k >= b->primary) --k;                 /* the sentinel is not stored */
    blk = b->bwt + (size_t)(k / BLK_BASES) * BLK_WORDS;
    return blk[c] + count_in_block(blk + 4, k % BLK_BASES + 1, c);
}

static int same_block(const orc_bwt_t *b, uint32_t k, uint32_t l)
{
    uint32_t kk = k >= b->primary ? k - 1 : k, ll = l >= b->primary ? l - 1 : l;
    return !(ll / BLK_BASES != kk / BLK_BASES || k == 0xffffffffu || l == 0xffffffffu);
}

/* bwt_2occ, libbwa/bwt.h:124-160: same values as two bwt_occ calls */
void orc_2occ(const orc_bwt_t *b, uint32_t k, uint32_t l, int c, uint32_t *ok, uint32_t *ol)
{
    uint64_t before = orc_blk_touches;
    *ok = orc_occ(b, k, c);
    *ol = (k == l) ? *ok : orc_occ(b, l, c);
    orc_blk_touches = before + ((k == l || same_block(b, k, l)) ? 1 : 2);
}

/* bwt_occ4, libbwa/bwt.h:165-182 (no seq_len special case there) */
void orc_occ4(const orc_bwt_t *b, uint32_t k, uint32_t cnt[4])
{
    const uint32_t *blk;
    int c;
    ++orc_blk_touches;
    if (k == 0xffffffffu) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    if (k >= b->primary) --k;
    blk = b->bwt + (size_t)(k / BLK_BASES) * BLK_WORDS;
    for (c = 0; c < 4; ++c) cnt[c] = blk[c] + count_in_block(blk + 4, k % BLK_BASES + 1, c);
}

/* bwt_2occ4, libbwa/bwt.h:185-222 */
void orc_2occ4(const orc_bwt_t *b, uint32_t k, uint32_t l, uint32_t ck[4], uint32_t cl[4])
{
    uint64_t before = orc_blk_touches;
    orc_occ4(b, k, ck);
    if (k == l) memcpy(cl, ck, 16); else orc_occ4(b, l, cl);
    orc_blk_touches = before + ((k == l || same_block(b, k, l)) ? 1 : 2);
}

/* bwt_B0 / bwt_invPsi / bwt_sa, libbwa/bwt.h:58-70, libbwa/bwt.c:69-79 */
static int stored_base(const orc_bwt_t *b, uint32_t k)
{
    const uint32_t *blk = b->bwt + (size_t)(k / BLK_BASES) * BLK_WORDS + 4;
    uint32_t j = k % BLK_BASES;
    return (int)((blk[j >> 4] >> ((15u - (j & 15u)) << 1)) & 3u);
}
uint32_t orc_sa(const orc_bwt_t *b, uint32_t k)
{
    uint32_t steps = 0;
    while (k % b->sa_intv != 0) {
        ++steps;
        if (k == b->primary) k = 0;
        else {
            int c = stored_base(b, k < b->primary ? k : k - 1);
            uint64_t keep = orc_blk_touches;      /* B0 shares its block with the occ that follows */
            k = b->L2[c] + orc_occ(b, k, c);
            orc_blk_touches = keep + 1;
        }
    }
    return steps + b->sa[k / b->sa_intv];
}

/* bwt_match_exact_alt, libbwa/bwt.c:102-117 */
int orc_match_exact_alt(const orc_bwt_t *b, int len, const uint8_t *str, uint32_t *k0, uint32_t *l0)
{
    uint32_t k = *k0, l = *l0, ok, ol;
    int i;
    for (i = len - 1; i >= 0; --i) {
        int c = str[i];
        if (c > 3) return 0;
        orc_2occ(b, k - 1, l, c, &ok, &ol);
        k = b->L2[c] + ok + 1;
        l = b->L2[c] + ol;
        if (k > l) return 0;
    }
    *k0 = k; *l0 = l;
    return (int)(l - k + 1);
}

/* bwa_cal_maxdiff, libbwa/bwtaln.c:58-70 (int x overflows exactly as there) */
int orc_cal_maxdiff(int l, double err, double thres)
{
    double elambda = exp(-l * err), sum = elambda, y = 1.0;
    int k, x = 1;
    for (k = 1; k < 1000; ++k) {
        y *= l * err;
        x = (int)((unsigned)x * (unsigned)k);
        sum += elambda * y / x;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

/* bwt_cal_width, libbwa/bwtaln.c:73-97 */
int orc_cal_width(const orc_bwt_t *b, int len, const uint8_t *str, orc_width_t *width)
{
    uint32_t k = 0, l = b->seq_len, ok, ol;
    int i, bid = 0;
    for (i = 0; i < len; ++i) {
        int c = str[i];
        if (c < 4) {
            orc_2occ(b, k - 1, l, c, &ok, &ol);
            k = b->L2[c] + ok + 1;
            l = b->L2[c] + ol;
        }
        if (k > l || c > 3) { k = 0; l = b->seq_len; ++bid; }
        width[i].w = l - k + 1;
        width[i].bid = bid;
    }
    width[len].w = 0;
    width[len].bid = ++bid;
    return bid;
}

/* ---- the score-bucketed stack (libbwa/bwtgap.c:13-91, bwtgap.h:7-22) ------ */

typedef struct {
    uint32_t k, l;
    int i, a, n_mm, n_gapo, n_gape, state, score;
    int last_diff_pos;         /* only written on diff pushes: slot memory persists (bwtgap.c:60) */
} orc_entry_t;
typedef struct { int n, cap; orc_entry_t *e; } orc_bucket_t;
struct orc_stack { int n_buckets, best, n_entries; orc_bucket_t *b; };

orc_stack_t *orc_stack_new(int n_buckets)
{
    orc_stack_t *s = (orc_stack_t *)calloc(1, sizeof(*s));
    int i;
    s->n_buckets = n_buckets;
    s->b = (orc_bucket_t *)calloc((size_t)n_buckets, sizeof(orc_bucket_t));
    for (i = 0; i < n_buckets; ++i) { s->b[i].cap = 4; s->b[i].e = (orc_entry_t *)calloc(4, sizeof(orc_entry_t)); }
    return s;
}
void orc_stack_free(orc_stack_t *s)
{
    int i;
    if (!s) return;
    for (i = 0; i < s->n_buckets; ++i) free(s->b[i].e);
    free(s->b); free(s);
}
static void stack_reset(orc_stack_t *s)
{
    int i;
    for (i = 0; i < s->n_buckets; ++i) s->b[i].n = 0;
    s->best = s->n_buckets; s->n_entries = 0;
}
static int score_of(const orc_gap_opt_t *o, int mm, int go, int ge) { return mm * o->s_mm + go * o->s_gapo + ge * o->s_gape; }

static void stack_push(orc_stack_t *s, const orc_gap_opt_t *o, int a, int i, uint32_t k, uint32_t l, int mm, int go, int ge,
                       int state, int is_diff)
{
    int sc = score_of(o, mm, go, ge);
    orc_bucket_t *q = s->b + sc;
    orc_entry_t *p;
    if (q->n == q->cap) {
        /* realloc in the reference leaves the new half uninitialised; every slot is written
         * by a diff push (or is a bucket-0 calloc slot) before its last_diff_pos is read */
        q->e = (orc_entry_t *)realloc(q->e, sizeof(orc_entry_t) * (size_t)(q->cap * 2));
        memset(q->e + q->cap, 0, sizeof(orc_entry_t) * (size_t)q->cap);
        q->cap *= 2;
    }
    p = q->e + q->n;
    p->k = k; p->l = l; p->i = i; p->a = a; p->n_mm = mm; p->n_gapo = go; p->n_gape = ge; p->state = state; p->score = sc;
    if (is_diff) p->last_diff_pos = i;
    ++q->n; ++s->n_entries;
    if (s->best > sc) s->best = sc;
}
static orc_entry_t stack_pop(orc_stack_t *s)
{
    orc_bucket_t *q = s->b + s->best;
    orc_entry_t e = q->e[--q->n];
    --s->n_entries;
    if (q->n == 0 && s->n_entries) {
        int i;
        for (i = s->best + 1; i < s->n_buckets; ++i) if (s->b[i].n) break;
        s->best = i;
    } else if (s->n_entries == 0) s->best = s->n_buckets;
    return e;
}

/* gap_shadow, libbwa/bwtgap.c:81-91 */
static void shadow(uint32_t x, uint32_t max, int last_diff_pos, orc_width_t *w)
{
    int i, j = 0;
    for (i = 0; i < last_diff_pos; ++i) {
        if (w[i].w > x) w[i].w -= x;
        else if (w[i].w == x) { w[i].bid = 1; w[i].w = max - (uint32_t)(++j); }
    }
}

static int ilog2(uint32_t v) { int c = 0; while (v >>= 1) ++c; return c; }

#define ST_M 0
#define ST_I 1
#define ST_D 2
#define MODE_GAPE 0x01
#define MODE_LOGGAP 0x04
#define MODE_NONSTOP 0x10

/* bwt_match_gap, libbwa/bwtgap.c:104-264 */
int orc_match_gap(const orc_bwt_t *const bwts[2], int len, const uint8_t *const seq[2], orc_width_t *const w[2],
                  orc_width_t *const seed_w[2], const orc_gap_opt_t *opt, orc_stack_t *stack, orc_aln_t *out, int cap)
{
    int best_score = score_of(opt, opt->max_diff + 1, opt->max_gapo + 1, opt->max_gape + 1);
    int best_diff = opt->max_diff + 1, max_diff = opt->max_diff, best_cnt = 0, n_aln = 0, j, n_N = 0;
    (void)best_diff;
    for (j = 0; j < len; ++j) if (seq[0][j] > 3) ++n_N;
    if (n_N > max_diff) return 0;

    stack_reset(stack);
    stack_push(stack, opt, 0, len, 0, bwts[0]->seq_len, 0, 0, 0, ST_M, 0);
    stack_push(stack, opt, 1, len, 0, bwts[0]->seq_len, 0, 0, 0, ST_M, 0);

    while (stack->n_entries) {
        orc_entry_t e;
        const orc_bwt_t *bwt;
        const uint8_t *str;
        orc_width_t *width;
        const orc_width_t *sw = 0;
        uint32_t k, l, ck[4], cl[4], occ;
        int a, i, m, m_seed = 0, hit, allow_diff, allow_M, gaps;

        if ((uint64_t)stack->n_entries > orc_peak_entries) orc_peak_entries = (uint64_t)stack->n_entries;
        if (stack->n_entries > opt->max_entries) break;
        e = stack_pop(stack);
        ++orc_pops;
        k = e.k; l = e.l; a = e.a; i = e.i;
        if (!(opt->mode & MODE_NONSTOP) && e.score > best_score + opt->s_mm) break;

        m = max_diff - (e.n_mm + e.n_gapo);
        if (opt->mode & MODE_GAPE) m -= e.n_gape;
        if (m < 0) continue;
        bwt = bwts[1 - a]; str = seq[a]; width = w[a];
        if (seed_w) {
            sw = seed_w[a];
            m_seed = opt->max_seed_diff - (e.n_mm + e.n_gapo);
            if (opt->mode & MODE_GAPE) m_seed -= e.n_gape;
        }
        if (i > 0 && m < width[i - 1].bid) continue;

        hit = 0;
        if (i == 0) hit = 1;
        else if (m == 0 && (e.state == ST_M || (opt->mode & MODE_GAPE) || e.n_gape == opt->max_gape)) {
            if (orc_match_exact_alt(bwt, i, str, &k, &l)) hit = 1;
            else continue;
        }
        if (hit) {
            int sc = score_of(opt, e.n_mm, e.n_gapo, e.n_gape), add = 1;
            if (n_aln == 0) {
                best_score = sc;
                best_diff = e.n_mm + e.n_gapo;
                if (opt->mode & MODE_GAPE) best_diff += e.n_gape;
                if (!(opt->mode & MODE_NONSTOP))
                    max_diff = (best_diff + 1 > opt->max_diff) ? opt->max_diff : best_diff + 1;
            }
            if (sc == best_score) best_cnt += (int)(l - k + 1);
            else if (best_cnt > opt->max_top2) break;
            if (e.n_gapo) {
                int n_cmp = n_aln < cap ? n_aln : cap;
                for (j = 0; j < n_cmp; ++j) if (out[j].k == k && out[j].l == l) break;
                if (j < n_cmp) add = 0;
            }
            if (add) {
                shadow(l - k + 1, bwt->seq_len, e.last_diff_pos, width);
                if (n_aln < cap) {
                    orc_aln_t *p = out + n_aln;
                    p->n_mm = (uint8_t)e.n_mm; p->n_gapo = (uint8_t)e.n_gapo; p->n_gape = (uint8_t)e.n_gape; p->a = (uint8_t)a;
                    p->k = k; p->l = l; p->score = sc;
                }
                ++n_aln;
            }
            continue;
        }

        --i;
        orc_2occ4(bwt, k - 1, l, ck, cl);
        occ = l - k + 1;
        allow_diff = allow_M = 1;
        if (i > 0) {
            int ii = i - (len - opt->seed_len);
            if (width[i - 1].bid > m - 1) allow_diff = 0;
            else if (width[i - 1].bid == m - 1 && width[i].bid == m - 1 && width[i - 1].w == width[i].w) allow_M = 0;
            if (seed_w && ii > 0) {
                if (sw[ii - 1].bid > m_seed - 1) allow_diff = 0;
                else if (sw[ii - 1].bid == m_seed - 1 && sw[ii].bid == m_seed - 1 && sw[ii - 1].w == sw[ii].w) allow_M = 0;
            }
        }
        gaps = (opt->mode & MODE_LOGGAP) ? ilog2((uint32_t)(e.n_gape + e.n_gapo)) / 2 + 1 : e.n_gapo + e.n_gape;
        if (allow_diff && i >= opt->indel_end_skip + gaps && len - i >= opt->indel_end_skip + gaps) {
            if (e.state == ST_M) {
                if (e.n_gapo < opt->max_gapo) {
                    stack_push(stack, opt, a, i, k, l, e.n_mm, e.n_gapo + 1, e.n_gape, ST_I, 1);
                    for (j = 0; j < 4; ++j) {
                        uint32_t kk = bwt->L2[j] + ck[j] + 1, ll = bwt->L2[j] + cl[j];
                        if (kk <= ll) stack_push(stack, opt, a, i + 1, kk, ll, e.n_mm, e.n_gapo + 1, e.n_gape, ST_D, 1);
                    }
                }
            } else if (e.state == ST_I) {
                if (e.n_gape < opt->max_gape)
                    stack_push(stack, opt, a, i, k, l, e.n_mm, e.n_gapo, e.n_gape + 1, ST_I, 1);
            } else {
                if (e.n_gape < opt->max_gape && (e.n_gape + e.n_gapo < max_diff || occ < (uint32_t)opt->max_del_occ)) {
                    for (j = 0; j < 4; ++j) {
                        uint32_t kk = bwt->L2[j] + ck[j] + 1, ll = bwt->L2[j] + cl[j];
                        if (kk <= ll) stack_push(stack, opt, a, i + 1, kk, ll, e.n_mm, e.n_gapo, e.n_gape + 1, ST_D, 1);
                    }
                }
            }
        }
        if (allow_diff && allow_M) {
            for (j = 1; j <= 4; ++j) {
                int c = (str[i] + j) & 3;
                int is_mm = (j != 4 || str[i] > 3);
                uint32_t kk = bwt->L2[c] + ck[c] + 1, ll = bwt->L2[c] + cl[c];
                if (kk <= ll) stack_push(stack, opt, a, i, kk, ll, e.n_mm + is_mm, e.n_gapo, e.n_gape, ST_M, is_mm);
            }
        } else if (str[i] < 4) {
            int c = str[i] & 3;
            uint32_t kk = bwt->L2[c] + ck[c] + 1, ll = bwt->L2[c] + cl[c];
            if (kk <= ll) stack_push(stack, opt, a, i, kk, ll, e.n_mm, e.n_gapo, e.n_gape, ST_M, 0);
        }
    }
    return n_aln;
}

/* per-read body of bwa_cal_sa_reg_gap, src/BwtMapper.cpp:104-137 */
int orc_align_read(const orc_bwt_t *const bwts[2], const uint8_t *fwd, int len, const orc_gap_opt_t *opt, double fnr,
                   int slice_max_gapo, orc_stack_t *stack, orc_aln_t *out, int cap)
{
    uint8_t s0[1024], s1[1024];
    orc_width_t w0[1025], w1[1025], sw0[64], sw1[64];
    const uint8_t *seq[2] = {s0, s1};
    orc_width_t *w[2] = {w0, w1}, *seedw[2] = {sw0, sw1};
    orc_gap_opt_t lo = *opt;
    int i;
    for (i = 0; i < len; ++i) {              /* seq = reversed read, rseq = reverse complement (BwtMapper.cpp:580-587) */
        s0[i] = fwd[len - 1 - i];
    }
    /* rseq: memcpy(seq) then seq_reverse(.., is_comp): position j holds comp(read[len-1-j]) */
    for (i = 0; i < len; ++i) s1[i] = s0[i] < 4 ? (uint8_t)(3 - s0[i]) : s0[i];
    orc_cal_width(bwts[0], len, s0, w0);
    orc_cal_width(bwts[1], len, s1, w1);
    if (fnr > 0.0) lo.max_diff = orc_cal_maxdiff(len, 0.02, fnr);
    lo.max_gapo = slice_max_gapo;
    lo.seed_len = opt->seed_len < len ? opt->seed_len : 0x7fffffff;
    if (len > opt->seed_len) {
        orc_cal_width(bwts[0], opt->seed_len, s0 + (len - opt->seed_len), sw0);
        orc_cal_width(bwts[1], opt->seed_len, s1 + (len - opt->seed_len), sw1);
    }
    return orc_match_gap(bwts, len, seq, w, len <= opt->seed_len ? 0 : seedw, &lo, stack, out, cap);
}

/* ---- read prep ------------------------------------------------------------ */

/* bwa_trim_read, libbwa/bwaseqio.c:75-88 (BWA_MIN_RDLEN 35) */
int orc_trim_len(int trim_qual, const uint8_t *qual, int len)
{
    int s = 0, l, max = 0, max_l = len - 1;
    if (trim_qual < 1 || qual == 0) return len;
    for (l = len - 1; l >= 35 - 1; --l) {
        s += trim_qual - ((int)qual[l] - 33);
        if (s < 0) break;
        if (s > max) { max = s; max_l = l; }
    }
    return max_l + 1;
}

/* IsReadInHashByCountMoreChunck + CountKmerHitInHash, src/BwtIndexer.cpp:441-456,261-313 */
int orc_kmer_pass(const uint8_t *const tables[6], const uint8_t *S, int thresh)
{
    int chunk, j, count = 0;
    for (chunk = 0; chunk < 3; ++chunk) {
        uint64_t kmer = 0;
        uint32_t x[6];
        for (j = 0; j < 32; ++j) kmer = (kmer << 2) | S[32 * chunk + j];
        x[0] = (uint32_t)(kmer >> 32);
        x[1] = (uint32_t)kmer;
        x[2] = (uint32_t)((kmer & 0xffff000000000000ull) >> 32) | (uint32_t)(kmer & 0xffff);
        x[3] = (uint32_t)((kmer & 0xffffffff0000ull) >> 16);
        x[4] = (uint32_t)((kmer & 0xffff000000000000ull) >> 32) | (uint32_t)((kmer & 0xffff0000ull) >> 16);
        x[5] = (uint32_t)((kmer & 0xffff00000000ull) >> 16) | (uint32_t)(kmer & 0xffff);
        for (j = 0; j < 6; ++j) count += (tables[j][x[j] >> 3] >> (x[j] & 7)) & 1;
        if (count >= thresh) return 1;
    }
    return 0;
}

/* ---- drand48 (glibc: X <- 0x5DEECE66D*X + 0xB mod 2^48; srand48: X = seed<<16 | 0x330E) ---- */
void orc_srand48(orc_rng_t *r, long seed) { r->x = ((uint64_t)(uint32_t)seed << 16) | 0x330Eull; r->n_calls = 0; }
double orc_drand48(orc_rng_t *r)
{
    r->x = (0x5DEECE66Dull * r->x + 0xBull) & 0xFFFFFFFFFFFFull;
    ++r->n_calls;
    return (double)r->x * (1.0 / 281474976710656.0);
}

/* bwa_aln2seq_core with set_main=1, n_multi=0: libbwa/bwase.c:19-46 */
void orc_aln2seq_main(int n_aln, const orc_aln_t *aln, orc_rng_t *rng, orc_se_t *s)
{
    int i, best;
    uint32_t cnt = 0;
    if (n_aln == 0) { s->type = 0; s->c1 = s->c2 = 0; return; }
    best = aln[0].score;
    for (i = 0; i < n_aln; ++i) {
        const orc_aln_t *p = aln + i;
        uint32_t width = p->l - p->k + 1;
        if (p->score > best) break;
        if (orc_drand48(rng) * (width + cnt) > (double)cnt) {
            s->n_mm = p->n_mm; s->n_gapo = p->n_gapo; s->n_gape = p->n_gape; s->strand = p->a; s->score = p->score;
            s->sa = p->k + (uint32_t)(width * orc_drand48(rng));
        }
        cnt += width;
    }
    s->c1 = cnt;
    for (; i < n_aln; ++i) cnt += aln[i].l - aln[i].k + 1;
    s->c2 = cnt - s->c1;
    s->type = s->c1 > 1 ? 2 : 1;
}

void orc_fill_log_n(int g[256]) { int i; g[0] = 0; for (i = 1; i < 256; ++i) g[i] = (int)(4.343 * log(i) + 0.5); }

/* bwa_approx_mapQ, libbwa/bwase.c:102-111 */
int orc_approx_mapq(const orc_se_t *s, int mm, const int g[256])
{
    int n;
    if (s->c1 == 0) return 23;
    if (s->c1 > 1) return 0;
    if (s->n_mm == mm) return 25;
    if (s->c2 == 0) return 37;
    n = s->c2 >= 255 ? 255 : (int)s->c2;
    return 23 < g[n] ? 0 : 23 - g[n];
}
