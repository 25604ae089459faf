/* fastquick_b200 -- C ABI of the B200-native FASTQuick align+summarize hot path.
 *
 * Plain pointers and sizes only; every call returns an int status (0 = ok,
 * negative = error, text via fqb_last_error()).  No exceptions cross this
 * boundary.  The reference has no FFI layer for this path (it is a C++ object
 * graph); each entry point below names the reference seam it replaces
 * (paths relative to the Griffan/FASTQuick tree).  INTEGRATION.md shows the
 * shim a maintainer adds on the reference side.
 */
#ifndef FASTQUICK_B200_H_
#define FASTQUICK_B200_H_
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FQB_OK            0
#define FQB_ERR_ARG      -1
#define FQB_ERR_IO       -2
#define FQB_ERR_CUDA     -3
#define FQB_ERR_STATE    -4
#define FQB_ERR_LIMIT    -5

#define FQB_MAX_READ_LEN 256      /* bwa_seq_t::len is 20 bits in the reference; reads on this path are <= ~250 bp */
#define FQB_BATCH_PAIRS  0x40000  /* READ_BUFFER_SIZE, src/BwtMapper.h:37 */

/* numeric fields of gap_opt_t (libbwa/bwtaln.h:98-119) that the path reads;
 * defaults = gap_init_opt() (libbwa/bwtaln.c:24-48) */
typedef struct {
    int32_t s_mm, s_gapo, s_gape;
    int32_t mode;
    int32_t indel_end_skip, max_del_occ, max_entries;
    double  fnr;
    int32_t max_diff, max_gapo, max_gape;
    int32_t max_seed_diff, seed_len;
    int32_t max_top2;
    int32_t trim_qual;
    int32_t flank_len, flank_long_len;
    int32_t read_len;
    int32_t kmer_thresh;            /* BwtIndexer(thresh), src/FASTQuick.cpp:174,362 */
    int32_t is_il13;                /* --I: qualities are phred+64 (BWA_MODE_IL13) */
} fqb_gap_opt_t;

/* pe_opt_t (libbwa/bwtaln.h:124-130); defaults = bwa_init_pe_opt() (libbwa/bwape.c:7-20) */
typedef struct {
    int32_t  max_isize, force_isize;
    uint32_t max_occ;
    int32_t  n_multi, N_multi;
    int32_t  type, is_sw;
    double   ap_prior;
} fqb_pe_opt_t;

void fqb_gap_opt_default(fqb_gap_opt_t *o);
void fqb_pe_opt_default(fqb_pe_opt_t *o);

/* one SA interval hit: bwt_aln1_t (libbwa/bwtaln.h:34-38) */
typedef struct {
    uint32_t k, l;
    int32_t  score;
    uint8_t  n_mm, n_gapo, n_gape, a;
} fqb_aln_t;

/* isize_info_t (libbwa/bwape.h:92-95) */
typedef struct {
    double   avg, std, ap_prior;
    uint32_t low, high, high_bayesian;
    uint32_t pad_;
} fqb_isize_t;

#define FQB_MAX_CIGAR 24
#define FQB_MAX_MULTI 11           /* n_multi/N_multi + 1, src/BwtMapper.cpp:857-871 */

/* per-read result row: the bwa_seq_t fields (libbwa/bwtaln.h:57-86) that
 * StatCollector::AddAlignment and BwtMapper::SetSamRecord consume */
typedef struct {
    uint32_t pos;                  /* pac coordinate */
    uint32_t sa;
    uint32_t c1, c2;
    int32_t  score;
    int32_t  len, full_len, clip_len;
    uint8_t  type, strand, filtered, extra_flag;
    uint8_t  n_mm, n_gapo, n_gape, mapQ;
    uint8_t  seQ, n_cigar, n_multi, has_cigar;
    uint16_t nm, n_aln;
    uint16_t cigar[FQB_MAX_CIGAR]; /* bwa_cigar_t: op<<14 | len */
} fqb_read_t;

typedef struct fqb_handle fqb_handle;

/* ---- index ------------------------------------------------------------ */
/* Replaces BwtIndexer::LoadIndex (src/BwtIndexer.cpp:803-837): reads
 * <prefix>.{bwt,rbwt,sa,rsa,pac,ann,amb} unchanged and uploads the re-laid tables to `device`
 * (CUDA ordinal).  The k-mer tables of ReadRollHashTable (src/BwtIndexer.cpp:569-579, 3 GiB in
 * <prefix>.rollhash) are rebuilt on the device from the flank text, bit for bit what
 * AddSeq2HashCore (611-713) produced; the file is only read when a flank name carries no "@x/y"
 * alleles or FQB_ROLLHASH_FROM_FILE is set in the environment. */
int fqb_create(const char *index_prefix, const fqb_gap_opt_t *gopt, const fqb_pe_opt_t *popt,
               int device, fqb_handle **out);
/* where the k-mer tables came from: 0 none (kmer_thresh == 0), 1 built on the device,
 * 2 uploaded from memory, 3 streamed from <prefix>.rollhash */
int fqb_kmer_tables_origin(const fqb_handle *h);
/* bytes [offset, offset + n_bytes) of the 6 x 2^29-byte tables, in .rollhash file order (tests) */
int fqb_kmer_tables_fetch(fqb_handle *h, uint64_t offset, uint64_t n_bytes, uint8_t *out);
void fqb_destroy(fqb_handle *h);
const char *fqb_last_error(void);

/* index facts the host shim needs (bwt_t / bntseq_t scalars) */
int fqb_index_info(const fqb_handle *h, int64_t *l_pac, int32_t *n_contigs,
                   uint32_t *primary /*[2]*/, uint32_t *seed);

/* ---- hot path, one batch of read pairs --------------------------------- */
/* Replaces the per-batch body of BwtMapper::PairEndMapper
 * (src/BwtMapper.cpp:1933-1982): bwa_read_seq_with_hash_dev's per-read prep
 * (476-613), bwa_cal_sa_reg_gap (63-168), bwa_cal_pac_pos_pe (721-907),
 * bwa_paired_sw (libbwa/bwape.c:463) and bwa_refine_gapped (libbwa/bwase.c:339).
 * bases/quals: ASCII, n_pairs rows of `stride` bytes per end (host memory,
 * pinned for full speed); lens: read lengths.  rows[e][i] receives the result
 * for end e of pair i.  Batches must be submitted in file order (the drand48
 * stream and last_ii carry over, src/BwtMapper.cpp:1817,780). */
int fqb_align_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride,
                    const uint8_t *bases1, const uint8_t *quals1, const int32_t *lens1,
                    const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2,
                    fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out);

/* ---- BAM output (SetSamFileHeader / SetSamRecord / BamIO.writeRecord, src/BwtMapper.cpp:947-1264, 2068-2074) ----
 * fqb_bam_open needs fqb_stats_open first (it loads <reference>.fai for the @SQ lines); rg_line is the --RG value
 * ("@RG\tID:foo\tSM:bar") or NULL.  fqb_bam_emit appends the records of the resident batch, after
 * fqb_stage_sw_refine and -- when statistics are collected -- after fqb_stage_stats, as the reference writes them
 * after StatCollector::AddAlignment; names as for fqb_stats_emit; bases/quals are the batch's host arrays. */
int fqb_bam_open(fqb_handle *h, const char *path, const char *rg_line);
int fqb_bam_emit(fqb_handle *h, const char *names, int32_t name_stride,
                 const uint8_t *bases1, const uint8_t *quals1, const uint8_t *bases2, const uint8_t *quals2, int32_t stride);
/* the same with the names of the SECOND reads given separately (names2 = NULL: same as names): mates whose FASTQ names differ
 * beyond a trailing /1 /2 keep their own names, as bwa_read_seq_with_hash_dev / SetSamRecord leave them */
int fqb_bam_emit2(fqb_handle *h, const char *names, const char *names2, int32_t name_stride,
                  const uint8_t *bases1, const uint8_t *quals1, const uint8_t *bases2, const uint8_t *quals2, int32_t stride);
/* sharded run inside one process (fqb_comm_init_local): the records handle h formats go, in call order, to owner's file */
int fqb_bam_attach(fqb_handle *h, fqb_handle *owner);
int fqb_bam_close(fqb_handle *h);

/* The reference's IO workers read batch n+1 while batch n is being mapped
 * (src/BwtMapper.cpp:1905-1931).  Same overlap here: upload the NEXT batch on a
 * copy stream into the second staging set; the following fqb_align_pairs /
 * fqb_stage_load with the same host pointers and shape uses it without copying.
 * The host buffers must stay untouched until that call. */
int fqb_prefetch_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride,
                       const uint8_t *bases1, const uint8_t *quals1, const int32_t *lens1,
                       const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2);

/* fqb_stage_load calls that found their batch already uploaded by fqb_prefetch_pairs (tests, tuning) */
uint64_t fqb_prefetch_hits(const fqb_handle *h);

/* ---- pipelined form of the per-batch body (src/BwtMapper.cpp:1905-1982: the reference reads batch n+1 on its IO
 * workers while batch n is mapped) ----------------------------------------------------------------------------
 * fqb_submit_pairs uploads a batch and enqueues its align stage (a1-a5: prep, k-mer filter, bwt_cal_width,
 * bwt_match_gap) on a second stream, into whichever of the handle's batch sets is free, and returns at once.
 * fqb_collect_pairs takes the OLDEST submitted batch through bwa_cal_pac_pos_pe, bwa_paired_sw, bwa_refine_gapped and,
 * when fqb_stats_open was called, StatCollector::AddAlignment's accumulation (= fqb_stage_stats), and starts the copy
 * of its result rows into rows1/rows2 (may be NULL); it does not wait.  Call order: submit(0); then for every n:
 * submit(n+1), collect(n).  At most two batches are in flight.  fqb_rows_wait blocks until the rows of the last
 * collected batch are complete and returns any error the device reported for the batches finished since the last
 * check; the stage-level fetch calls and fqb_stats_emit / fqb_bam_emit work on the batch collected last.
 * on_device != 0: the pointers are device pointers (see fqb_stage_load). */
int fqb_submit_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride,
                     const uint8_t *bases1, const uint8_t *quals1, const int32_t *lens1,
                     const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device);
int fqb_collect_pairs(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2);

/* ---- packed input form ------------------------------------------------------------------------------------------
 * What bwa_read_seq_with_hash_dev (src/BwtMapper.cpp:543-590) produces per read before the aligner sees it is the nt4
 * code of every base (nst_nt4_table, libbwa/bntseq.c:38-55) and its quality.  A batch can be handed over in that
 * form directly, which is 264 instead of 400 bytes per 2 x 100 bp pair over PCIe and lets prep_kernel fetch a read's
 * bases with 128-bit loads: `packed` rows hold 2 bits per base, 16 bases per little-endian 32-bit word from bit 0 up,
 * fqb_packed_stride(stride) = 16 * ceil(stride / 64) bytes per read; the quality rows (`stride` bytes per read, as
 * before) carry in bit 7 the flag "not A/C/G/T", and the 2-bit field of such a base holds its nt4 code minus 4 (0 for N
 * and every other letter, 1 for '-').  fqb_pack_reads converts `n` ASCII rows (FQB_ERR_ARG on a quality byte above
 * 127); fqb_feeder_fill_packed has the feeder's parse workers emit both forms.  The packed calls take the same
 * lens / on_device arguments as their ASCII counterparts and give bit-identical results. */
int32_t fqb_packed_stride(int32_t stride);
int fqb_pack_reads(int64_t n, int32_t stride, const uint8_t *bases, const uint8_t *quals,
                   int32_t packed_stride, uint8_t *packed_out, uint8_t *quals_out);
int fqb_align_pairs_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride,
                           const uint8_t *packed1, const uint8_t *quals1, const int32_t *lens1,
                           const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2,
                           fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out);
int fqb_stage_load_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride,
                          const uint8_t *packed1, const uint8_t *quals1, const int32_t *lens1,
                          const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2, int on_device);
int fqb_submit_pairs_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride,
                            const uint8_t *packed1, const uint8_t *quals1, const int32_t *lens1,
                            const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2, int on_device);

/* ---- stage-level entry points (parity tests, bench, profiling) ---------------
 * The same kernels fqb_align_pairs sequences, one group at a time, on the batch
 * made resident by fqb_stage_load.  Read index r = 2*pair + end. */
/* on_device != 0: the pointers are device pointers on the handle's GPU (inputs stay where they are) */
int fqb_stage_load(fqb_handle *h, int32_t n_pairs, int32_t stride,
                   const uint8_t *bases1, const uint8_t *quals1, const int32_t *lens1,
                   const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device);
/* a1-a5: read prep + k-mer filter (src/BwtMapper.cpp:543-590, src/BwtIndexer.cpp:524),
 * bwt_cal_width (libbwa/bwtaln.c:73) and bwt_match_gap (libbwa/bwtgap.c:104) = bwa_cal_sa_reg_gap */
int fqb_stage_align(fqb_handle *h);
/* a6-a9: bwa_cal_pac_pos_pe (src/BwtMapper.cpp:721-907) = bwa_aln2seq (drand48) + bwt_sa + bwa_approx_mapQ,
 * infer_isize (libbwa/bwape.c:49), pairing (libbwa/bwape.c:119) and the multi-hit counts */
int fqb_stage_pair(fqb_handle *h);
/* a10 + a11: bwa_paired_sw (libbwa/bwape.c:463-625) and bwa_refine_gapped incl. NM and bwa_correct_trimmed
 * (libbwa/bwase.c:183-418) */
int fqb_stage_sw_refine(fqb_handle *h);
/* ---- a12-a14: StatCollector ----------------------------------------------------------------
 * fqb_stats_open      = StatCollector::RestoreVcfSites + SetGenomeSize (src/StatCollector.cpp:1742, src/BwtMapper.cpp:225-226)
 * fqb_stats_begin_file = FileStatCollector FSC(fq1, fq2) per FASTQ pair (src/BwtMapper.cpp:249); restarts drand48 / last_ii
 * fqb_stage_stats     = StatCollector::AddAlignment for every pair of the batch (src/StatCollector.cpp:950-1101): pair
 *                       classification, insert-size bookkeeping, PCR-duplicate set, X/Y counters, per-base pile-up / depth /
 *                       quality / cycle accumulation (424-621) -- all on the device
 * fqb_stats_emit      = the InsertSizeTable lines of the batch (text; host)
 * fqb_stats_finish    = StatCollector::ProcessCore (2012-2028): writes <out_prefix>.{DepthDist,GCDist,EmpRepDist,EmpCycleDist,
 *                       AdjustedInsertSizeDist,RawInsertSizeDist,SexChromInfo,Pileup,FASTQ.csv,Sequence.csv,Summary,vcf} */
/* --targetRegion: StatCollector::SetTargetRegion (src/StatCollector.cpp:2284-2288); call before fqb_stats_open */
int fqb_stats_set_target_region(fqb_handle *h, const char *bed_path);
int fqb_stats_open(fqb_handle *h, const char *index_prefix);
int fqb_stats_begin_file(fqb_handle *h, const char *out_prefix, const char *fastq1, const char *fastq2);
/* counters of the current file so far, as the reference prints them after each file (src/BwtMapper.cpp:2116-2122):
 * out6 = TotalFiltered, BwaUnmapped (pairs; printed x 2), TotalMAPQ, TotalRetained, NumBase, NumRead */
int fqb_stats_file_counters(fqb_handle *h, int64_t *out6);
/* every accumulator back to zero (the state right after fqb_stats_open): a new run on the same handle */
int fqb_stats_reset(fqb_handle *h);
int fqb_stage_stats(fqb_handle *h);
int fqb_stats_emit(fqb_handle *h, const char *names, int32_t name_stride);
/* names2: names of the second reads (NULL = same as names); a line for a pair whose first read is unmapped carries the second
 * read's name (q->name, src/StatCollector.cpp:695,708) */
int fqb_stats_emit2(fqb_handle *h, const char *names, const char *names2, int32_t name_stride);
/* fqb_stats_emit and fqb_bam_emit copy what they need from the device, then format and write.  With FQB_ASYNC_EMIT=1 in the
 * environment that host phase runs on threads of its own while the caller submits the next batch (off by default: it only
 * pays on hosts with idle cores); the host arrays passed to the two calls (names, bases, quals) must then stay untouched
 * until the next fqb_stats_emit on the handle returns (it joins both), or fqb_emit_sync / fqb_stats_finish /
 * fqb_stats_close_table / fqb_bam_close / fqb_stats_begin_file does, and a write error met by those threads is returned by
 * the call that joins them.  fqb_emit_sync is a no-op otherwise. */
int fqb_emit_sync(fqb_handle *h);
int fqb_stats_finish(fqb_handle *h, const char *out_prefix);
/* InsertSizeEstimator (src/InsertSizeEstimator.cpp:43-173) alone, host only: <table_path> is a finished InsertSizeTable,
 * <out_path> receives what fqb_stats_finish writes as <prefix>.AdjustedInsertSizeDist. */
int fqb_isize_adjusted_file(const char *table_path, const char *out_path);
/* The host half of infer_isize (libbwa/bwape.c:49-117) on its own: hist[v] = number of pairs with both mapQ >= 20 and
 * insert size v < 100000 (100,000 bins; the pair stage collects it on the device), max_len = the batch's longest read,
 * L = the BWT's seq_len.  Returns 1 and fills *ii, or 0 when inference fails (fewer than 20 pairs, degenerate spread)
 * and *ii is "unset" (avg = std = -1).  fqb_isize_penalty: the table pairing() reads, entry l =
 * (int)(-4.343 * log(.5 * erfc(M_SQRT1_2 * fabs(l - avg) / std)) + .499) (libbwa/bwape.h:62) for l = 0..high_bayesian;
 * returns the table's length and writes at most cap entries. */
int fqb_infer_isize_hist(const uint32_t *hist, int32_t max_len, double ap_prior, int64_t L, fqb_isize_t *ii);
int64_t fqb_isize_penalty(const fqb_isize_t *ii, int32_t *out, int64_t cap);
/* The libm-dependent tables fqb_create evaluates on the host and keeps on the device: maxdiff[l] = bwa_cal_maxdiff(l, 0.02,
 * gopt->fnr) (libbwa/bwtaln.c:58-70; gopt->max_diff when fnr <= 0) for l = 0..FQB_MAX_READ_LEN, and g_log_n
 * (libbwa/bwase.c:602-606), 256 entries. */
int fqb_host_tables(const fqb_gap_opt_t *gopt, int32_t *maxdiff, int32_t *log_n);
/* n_stacks of gap_init_stack (libbwa/bwtgap.c:18) for a batch whose longest read has max_len bases, after the max_gapo
 * clamp of src/BwtMapper.cpp:73-81: (max_diff + 1) * s_mm + (max_gapo + 1) * s_gapo + (max_gape + 1) * s_gape. */
int fqb_search_buckets(const fqb_gap_opt_t *gopt, int32_t max_len);
/* Sharded runs (one handle per GPU, batches dealt round-robin): each handle writes the InsertSizeTable lines
 * (StatCollector::AddAlignment's `fout`, src/StatCollector.cpp:950) of its own batches.  fqb_stats_close_table
 * finishes a handle's file; fqb_stats_merge_tables, on the handle that will call fqb_stats_finish, splices the
 * batches of the other handles' files (<other_prefix>.InsertSizeTable) back into file order. */
int fqb_stats_close_table(fqb_handle *h);
int fqb_stats_merge_tables(fqb_handle *h, const char *const *other_prefixes, int32_t n_others);
/* ---- multi-GPU (one process per GPU; reads shard by 262,144-pair batch, index replicated) -----------------
 * The only cross-batch state of the path is the drand48 position (srand48 once per FASTQ pair,
 * src/BwtMapper.cpp:1817) and last_ii (src/BwtMapper.cpp:780): the owner of batch b hands both to the owner of
 * batch b+1 between fqb_stage_align and fqb_stage_pair.  At the end the integer accumulators are combined with
 * an NCCL reduce (groups 0-2: sum; group 3: min) on buffers moved with fqb_stats_export / fqb_stats_import. */
int fqb_get_stream_state(fqb_handle *h, uint64_t *rng_calls, fqb_isize_t *last_ii);
/* The one draw this library does not reproduce: bwa_aln2seq_core (libbwa/bwase.c:33-36) skips a read's only best interval
 * when drand48() returns exactly 0.0 and then takes one draw instead of two; the device counts two draws for such reads
 * and reports FQB_ERR_LIMIT should it ever meet that draw.  drand48 is a full-period 48-bit LCG, so after srand48(seed) it
 * returns 0.0 exactly once per 2^48 draws: this returns the 1-based number of that draw (host arithmetic, no device needed).
 * For the seed bns->seed = 11 of every BWA index it is draw 79,023,531,276,618 -- the stream restarts with every FASTQ pair
 * (src/BwtMapper.cpp:1817), so a single input file would have to hold about 2e13 pairs to reach it. */
uint64_t fqb_drand48_zero_index(uint32_t seed);
int fqb_set_stream_state(fqb_handle *h, uint64_t rng_calls, const fqb_isize_t *last_ii);
int fqb_set_pair_base(fqb_handle *h, uint64_t first_pair);
int fqb_stats_group_bytes(fqb_handle *h, int which, uint64_t *bytes);
int fqb_stats_export(fqb_handle *h, int which, void *dst_device);
int fqb_stats_import(fqb_handle *h, int which, const void *src_device);
/* Variable-size statistics state of a sharded run (host OR device buffers): which = 0 the marker pile-up entries
 * (20 bytes each; they carry their global pair index, so the merged pile-up keeps file order), which = 1 the
 * distinct PCR-duplicate keys of StatCollector's duplicateTable (8 bytes each; a key two handles both hold is
 * one more duplicated pair).  Import on the handle that will call fqb_stats_finish, after fqb_stats_import. */
int fqb_stats_var_count(fqb_handle *h, int which, uint64_t *n);
int fqb_stats_var_export(fqb_handle *h, int which, void *dst, uint64_t cap);
int fqb_stats_var_import(fqb_handle *h, int which, const void *src, uint64_t n);
/* ---- multi-GPU in the library itself (host side stays C/C++: the loop to shard is src/BwtMapper.cpp:1796-2143, batch unit
 * src/BwtMapper.h:36-37) -------------------------------------------------------------------------------------------
 * Global batch b of a file belongs to rank b % world.  fqb_comm_ring_handle creates the handle's mailbox for the
 * 56-byte hand-off (drand48 position + last_ii) and returns its cudaIpcMemHandle_t; fqb_comm_unique_id is
 * ncclGetUniqueId.  One process per GPU: the launcher distributes rank 0's id and every rank's mailbox handle, then each
 * rank calls fqb_comm_init (NCCL communicator for the end-of-run merge; the next rank's mailbox is mapped through CUDA
 * IPC and written over NVLink by a one-thread kernel).  One process driving several GPUs: fqb_comm_init_local on all its
 * handles (peer access instead of IPC).  fqb_collect_pairs_sharded is fqb_collect_pairs with the hand-off in the stream:
 * global_batch = position of the batch in file order, first_pair = global index of its first pair, is_last = no batch
 * follows in this file.  fqb_comm_merge_stats (every rank, after its last batch; terminal for the run): the distinct
 * PCR-duplicate keys go all-to-all to owner ranks chosen by hash (grouped ncclSend / ncclRecv, exact sizes) and every owner
 * counts the keys more than one rank holds (a key in m ranks is m - 1 more duplicated pairs); one grouped ncclReduce brings
 * the accumulators onto rank 0 (sum; first-touch contig order: min); the marker pile-up entries follow with exact-size
 * sends.  *ms_out = device time of the exchange.  Rank 0 then splices the ranks' InsertSizeTable batches
 * (fqb_stats_merge_tables) and writes the files (fqb_stats_finish). */
int fqb_comm_ring_handle(fqb_handle *h, uint8_t *out64);
int fqb_comm_unique_id(uint8_t *out128);
int fqb_comm_init(fqb_handle *h, int rank, int world, const uint8_t *nccl_id128, const uint8_t *ring_handles /* world x 64 */);
int fqb_comm_init_local(fqb_handle **hs, int n);
int fqb_collect_pairs_sharded(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2, uint64_t global_batch, uint64_t first_pair, int is_last);
int fqb_comm_merge_stats(fqb_handle *h, double *ms_out);
int fqb_stage_fetch_rows(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out);
/* asynchronous variant: the copies run on a side stream while the next batch is processed; fqb_rows_wait blocks until
 * rows1/rows2 of the last call are complete (pinned destination buffers for real overlap) */
int fqb_stage_fetch_rows_async(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2);
int fqb_rows_wait(fqb_handle *h);
/* new FASTQ pair: restart the drand48 stream and forget last_ii (src/BwtMapper.cpp:1811-1817).  In a sharded run every rank
 * calls it (or fqb_stats_begin_file, which does) once per file: it also starts a new epoch of the hand-off ring's sequence numbers */
int fqb_reset_stream(fqb_handle *h);
int fqb_stage_fetch_prep(fqb_handle *h, int32_t *len, int32_t *full_len, uint8_t *filtered,
                         uint8_t *codes, int32_t codes_stride);
int fqb_stage_fetch_aln(fqb_handle *h, int32_t cap, fqb_aln_t *out, int32_t *n_aln);
/* since creation: stack pops, rank-query pairs, reference-equivalent occ-block touches (N_blk, the
 * roofline's algorithmic unit), overflow reads of the last batch */
int fqb_stage_counters(fqb_handle *h, uint64_t *out4);
/* ---- host feeder (row f2) ------------------------------------------------
 * FASTQ file -> the fixed-stride batches fqb_align_pairs / fqb_prefetch_pairs take, decoded in parallel.
 * Replaces the record-by-record kseq_read3_fpc (libbwa/kseq.h:327-370) loop of bwa_read_seq_with_hash_dev
 * (src/BwtMapper.cpp:476-613): a producer thread turns the file into text blocks (BGZF members are inflated
 * independently by a worker pool; any other gzip stream is memory-mapped and decoded serially by the library's
 * own inflate loop, zlib's when the file cannot be mapped or FQB_GZIP_ZLIB is set; plain text is read as is) and the
 * pool parses runs of whole records straight into the caller's (pinned) batch.  n_threads <= 0 = one per
 * host core, at most 16; the pool is shared by all feeders of the process.
 * fill: up to n_max records; bases padded with 'N' and qualities with '!' to `stride`; names cut at the first
 * blank with a trailing /1 or /2 removed, zero-padded to name_stride.  Returns the number of records
 * (0 = end of file) or -1 (fqb_last_error: malformed record, read longer than stride, corrupt stream). */
typedef struct fqb_feeder fqb_feeder;
int fqb_feeder_open(const char *path, int n_threads, fqb_feeder **out);
int fqb_feeder_format(const fqb_feeder *f);         /* 0 plain text, 1 gzip stream, 2 BGZF */
int64_t fqb_feeder_fill(fqb_feeder *f, int32_t n_max, int32_t stride, uint8_t *bases, uint8_t *quals,
                        int32_t *lens, char *names, int32_t name_stride);
/* fill, plus the same rows in the packed upload form (see fqb_pack_reads): the workers that parse a record also pack it */
int64_t fqb_feeder_fill_packed(fqb_feeder *f, int32_t n_max, int32_t stride, uint8_t *bases, uint8_t *quals,
                               int32_t *lens, char *names, int32_t name_stride,
                               int32_t packed_stride, uint8_t *packed, uint8_t *quals_flagged);
void fqb_feeder_close(fqb_feeder *f);
/* The feeder's gzip-stream path on a buffer: all members of gz[0..n_gz) decoded into out[0..cap), through text
 * blocks of block_bytes (<= 0: the feeder's 4 MiB) exactly as the feeder chains them; *n_out = bytes written.
 * FQB_ERR_IO on a corrupt stream (CRC-32 / length / code checks), FQB_ERR_ARG if cap is too small. */
int fqb_gunzip(const uint8_t *gz, int64_t n_gz, uint8_t *out, int64_t cap, int32_t block_bytes, int64_t *n_out);
/* The BAM writer's member format on a buffer: data[0..n) cut into 0xff00-byte payloads, each compressed into one BGZF
 * member exactly as fqb_bam_emit's writer does (the library's own deflate, or zlib at FQB_BAM_LEVEL), members back to
 * back in out[0..cap), no end-of-file member; *n_out = bytes written.  cap >= n + 64 * (n / 0xff00 + 1) always fits. */
int fqb_bgzf_compress(const uint8_t *data, int64_t n, uint8_t *out, int64_t cap, int64_t *n_out);
/* The BAM writer's file layer on a buffer: data[0..n) handed to it in pieces of `piece` bytes (copied when owned == 0,
 * as whole-record chunks that start their own members otherwise), compressed by its writer thread and helpers, closed
 * with the BGZF end-of-file member. */
int fqb_bgzf_write_file(const char *path, const uint8_t *data, int64_t n, int64_t piece, int32_t owned);
void *fqb_host_alloc(size_t bytes);                /* pinned host memory for the feeder's batches */
void fqb_host_free(void *p);
uint64_t fqb_launch_count(const fqb_handle *h);   /* kernels launched by this handle so far */
/* accumulated device time (ms, CUDA events on the handle's stream) of the rank-query kernels of
 * fqb_stage_align (bwt_cal_width + bwt_match_gap fast pass) and the number of batches covered */
int fqb_rank_query_time(const fqb_handle *h, double *ms, uint64_t *launches);
void *fqb_stream(fqb_handle *h);   /* the cudaStream_t the handle launches on (for event timing) */
/* BW_L2, the roofline denominator of the rank-query kernels (occ primitives libbwa/bwt.h:89-222): every
 * thread of a grid that fills all SMs issues independent 256-bit loads at pseudo-random unit_bytes-aligned offsets
 * (unit_bytes = 32, 64 or 128 contiguous bytes per access, L1 bypassed) of a buffer_bytes buffer (8 MiB: L2-resident
 * like the FM index); best and median GB/s of `reps` launches after two warm-up launches. */
int fqb_measure_l2(int device, int64_t buffer_bytes, int32_t unit_bytes, int32_t reps, double *gbs_best, double *gbs_median);

/* ---- synthetic fixtures (bench + tests; hs37d5/dbSNP are not available offline) ----
 * Not part of the drop-in surface: these stand in for `FASTQuick index` output
 * (src/FASTQuick.cpp:38-157) and for FASTQ input so the hot path can be driven
 * where the reference binary is absent. */
typedef struct fqb_synth fqb_synth;
typedef struct {
    uint64_t seed;
    int32_t n_long, n_short, n_x, n_y;
    int32_t flank_short, flank_long, spacing;
    int32_t n_dup;       /* 200-bp segments copied from one marker's flank into another's: exact repeats in the index (default 0) */
} fqb_synth_ref_cfg_t;
typedef struct {
    uint64_t seed;
    int32_t read_len;
    int32_t max_indel_len;
    double f_on, sub_rate, ins_rate, del_rate, n_rate, isize_mean, isize_sd, bad_tail_rate;
} fqb_synth_read_cfg_t;
void fqb_synth_ref_cfg_default(fqb_synth_ref_cfg_t *c);
void fqb_synth_read_cfg_default(fqb_synth_read_cfg_t *c);
int fqb_synth_create(const fqb_synth_ref_cfg_t *cfg, fqb_synth **out);
void fqb_synth_destroy(fqb_synth *s);
/* genome.fa(+.fai,.amb), markers.vcf, dbsnp.vcf: the inputs of `FASTQuick index --predefinedVCF` */
int fqb_synth_write_inputs(const fqb_synth *s, const char *dir);
/* <prefix> = "<out_prefix>.FASTQuick.fa": all files BwtIndexer::BuildIndex + runIndex would leave */
int fqb_synth_write_index(const fqb_synth *s, const char *genome_path, const char *dbsnp_path,
                          const char *prefix, int with_rollhash);
/* fixture side: the BwtIndexer::BuildIndex files for an arbitrary flank FASTA (">chr:pos@R/A" names, one line per
 * sequence, as <prefix> itself is laid out); used to test flanks with N, lower case and '-' */
int fqb_index_from_flank_fasta(const char *flank_fasta, const char *prefix, int with_rollhash);
/* engine over the synthetic index built in memory (no files written) */
int fqb_create_from_synth(const fqb_synth *s, const fqb_gap_opt_t *gopt, const fqb_pe_opt_t *popt,
                          int device, fqb_handle **out);
int fqb_synth_reads(const fqb_synth *s, const fqb_synth_read_cfg_t *cfg, int64_t first_pair, int64_t n_pairs,
                    uint8_t *bases1, uint8_t *quals1, uint8_t *bases2, uint8_t *quals2, int n_threads);
/* gz FASTQ with fixed-width names r%011lld/1 and /2 */
int fqb_write_fastq_gz(const char *path, int which_end, int64_t first_pair, int64_t n_pairs, int32_t read_len,
                       const uint8_t *bases, const uint8_t *quals);

#ifdef __cplusplus
}
#endif
#endif
