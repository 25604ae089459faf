/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * CPU restatement (plain C) of the reference algorithm on the align hot path,
 * written from the reference's behaviour; each function cites the reference
 * file:line it follows.  Parity is PINNED: tests/test_oracle_vs_ref.py checks
 * every function here against the reference's own code compiled into
 * oracle/_ref/libfqref.so (recipe: oracle/Makefile.ref) on seeded inputs, and
 * against the golden vectors under tests/golden/.
 */
#ifndef FQ_ORACLE_H_
#define FQ_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t primary, L2[5], seq_len;
    const uint32_t *bwt;     /* reference interleaved layout: 4 counts + 8 words per 128 bases */
    uint32_t sa_intv, n_sa;
    const uint32_t *sa;      /* sa[0] = 0xffffffff */
} orc_bwt_t;

typedef struct { uint32_t w; int32_t bid; } orc_width_t;

typedef struct {
    int s_mm, s_gapo, s_gape, mode;
    int indel_end_skip, max_del_occ, max_entries;
    int max_diff, max_gapo, max_gape, max_seed_diff, seed_len, max_top2;
} orc_gap_opt_t;

typedef struct { uint32_t k, l; int32_t score; uint8_t n_mm, n_gapo, n_gape, a; } orc_aln_t;

/* counts occ-block touches as SURVEY.md §8(d) defines N_blk */
extern uint64_t orc_blk_touches;
extern uint64_t orc_pops, orc_peak_entries;

uint32_t orc_occ(const orc_bwt_t *b, uint32_t k, int c);
void orc_2occ(const orc_bwt_t *b, uint32_t k, uint32_t l, int c, uint32_t *ok, uint32_t *ol);
void orc_occ4(const orc_bwt_t *b, uint32_t k, uint32_t cnt[4]);
void orc_2occ4(const orc_bwt_t *b, uint32_t k, uint32_t l, uint32_t ck[4], uint32_t cl[4]);
uint32_t orc_sa(const orc_bwt_t *b, uint32_t k);
int orc_match_exact_alt(const orc_bwt_t *b, int len, const uint8_t *str, uint32_t *k0, uint32_t *l0);

int orc_cal_maxdiff(int l, double err, double thres);
int orc_cal_width(const orc_bwt_t *b, int len, const uint8_t *str, orc_width_t *width);

typedef struct orc_stack orc_stack_t;
orc_stack_t *orc_stack_new(int n_buckets);
void orc_stack_free(orc_stack_t *s);
/* returns n_aln; hits written to out (up to cap; count keeps growing past cap) */
int orc_match_gap(const orc_bwt_t *const bwts[2], int len, const uint8_t *const seq[2], orc_width_t *const w[2],
                  orc_width_t *const seed_w[2], const orc_gap_opt_t *opt, orc_stack_t *stack, orc_aln_t *out, int cap);

/* whole per-read driver = body of bwa_cal_sa_reg_gap; fwd = nt4 codes in read orientation */
int orc_align_read(const orc_bwt_t *const bwts[2], const uint8_t *fwd, int len, const orc_gap_opt_t *opt, double fnr,
                   int slice_max_gapo, orc_stack_t *stack, orc_aln_t *out, int cap);

/* read prep */
int orc_trim_len(int trim_qual, const uint8_t *qual, int len);
int orc_kmer_pass(const uint8_t *const tables[6], const uint8_t *fwd_codes, int thresh);

/* drand48 stream + hit selection */
typedef struct { uint64_t x; uint64_t n_calls; } orc_rng_t;
void orc_srand48(orc_rng_t *r, long seed);
double orc_drand48(orc_rng_t *r);
typedef struct { uint32_t sa, c1, c2; int32_t score; uint8_t n_mm, n_gapo, n_gape, strand, type; } orc_se_t;
void orc_aln2seq_main(int n_aln, const orc_aln_t *aln, orc_rng_t *rng, orc_se_t *s);
int orc_approx_mapq(const orc_se_t *s, int mm, const int g_log_n[256]);
void orc_fill_log_n(int g_log_n[256]);

/* ---- paired-end resolution: bwa_cal_pac_pos_pe (src/BwtMapper.cpp:721-907) ---------------- */
typedef struct { double avg, std, ap_prior; uint32_t low, high, high_bayesian, pad_; } orc_isize_t;
typedef struct {
    int32_t max_isize, force_isize; uint32_t max_occ; int32_t n_multi, N_multi, type, is_sw; double ap_prior;
} orc_pe_opt_t;
/* per-read row; same layout as fqb_read_t (include/fastquick_b200.h) so tests can compare bytes */
typedef struct {
    uint32_t pos, sa, c1, c2; int32_t score, len, full_len, clip_len;
    uint8_t type, strand, filtered, extra_flag, n_mm, n_gapo, n_gape, mapQ, seQ, n_cigar, n_multi, has_cigar;
    uint16_t nm, n_aln; uint16_t cigar[24];
} orc_row_t;
/* infer_isize (libbwa/bwape.c:49-117); returns 0 ok / -1 failed */
int orc_infer_isize(int n_pairs, const orc_row_t *rows /*2n, r=2*pair+end*/, orc_isize_t *ii, double ap_prior, int64_t L);
/* whole stage for one batch.  rows: in = len/full_len/clip_len/filtered set, rest zero; out = after bwa_cal_pac_pos_pe.
 * multi_pos (may be null): [2n][11] pac positions of the multi hits. */
int orc_cal_pac_pos_pe(const orc_bwt_t *const bwts[2], int n_pairs, orc_row_t *rows, const int32_t *n_aln,
                       const orc_aln_t *aln, int aln_cap, double fnr, int opt_max_diff, int s_mm,
                       const orc_pe_opt_t *popt, orc_rng_t *rng, const orc_isize_t *last_ii, orc_isize_t *ii,
                       uint32_t *multi_pos);

/* ---- banded global alignment + gapped refinement (libbwa/stdaln.c:345-524, libbwa/bwase.c:183-337) ---- */
int orc_global_align(const uint8_t *seq1, int len1, const uint8_t *seq2, int len2, int gap_open, int gap_ext, int gap_end,
                     int band, uint8_t *ops_out, int *n_ops);
int orc_ops_to_cigar(const uint8_t *ops, int n_ops, uint16_t *cigar);
int orc_refine_gapped(int64_t l_pac, const uint8_t *pac, int len, const uint8_t *seq, uint32_t *pos_io, int ext, uint16_t *cigar);
int orc_cal_nm(int n_cigar, const uint16_t *cigar, int has_cigar, int len, uint32_t pos, const uint8_t *seq, int64_t l_pac, const uint8_t *pac);
void orc_paired_sw(int64_t l_pac, const uint8_t *pac, int n_pairs, orc_row_t *rows, const uint8_t *codes, int stride,
                   const orc_pe_opt_t *popt, const orc_isize_t *ii);
/* RegionList of the reference for one chromosome (src/RegionList.cpp): arrays of (start, end) pairs sorted by start */
int orc_regions_add(int *r, int n, int start, int end);
int orc_regions_collapse(int *r, int n, long long *len);
int orc_regions_join(int *a, int na, const int *b, int nb, long long *len);
int orc_regions_overlapped(const int *r, int n, int pos);
/* MD tag (bwa_cal_md1, libbwa/bwase.c:234-296) and StatCollector::RecoverRefseqByMDandCigar (src/StatCollector.cpp:101-172) */
int orc_cal_md(int n_cigar, const uint16_t *cigar, int has_cigar, int len, uint32_t pos, const uint8_t *seq, int64_t l_pac,
               const uint8_t *pac, char *out, int cap);
int orc_recover_refseq(const char *read, const char *md, const uint16_t *cigar, int n_cigar, char *out, int cap);
void orc_refine_gapped_batch(int64_t l_pac, const uint8_t *pac, int n_reads, orc_row_t *rows, const uint8_t *codes, int stride);

#ifdef __cplusplus
}
#endif
#endif
