// Engine: owns the device-resident index and the per-batch device buffers, and
// sequences the kernels of the hot path on one CUDA stream.  Exposed through the C
// ABI in include/fastquick_b200.h.  One engine per GPU (one process per GPU).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fastquick_b200.h"
#include "fq_common.h"
#include "fq_hostmath.h"
#include "fq_index.h"
#include "fq_kernels.cuh"
#include "fq_kmer.cuh"
#include "fq_pair_kernels.cuh"
#include "fq_dp_kernels.cuh"
#include "fq_stats_kernels.cuh"
#include "fq_stats_host.h"
#include "fq_bam.h"
#include <algorithm>
#include <chrono>
#include <fstream>
#include <future>
#include <mutex>
#include <thread>
#include <dlfcn.h>
#include <nccl.h>
#include <cmath>
#include "fq_relayout.h"
#include "fq_synth.h"

std::vector<fqb::FlankSeq> fqb_synth_flanks_internal(const fqb_synth *s);

using namespace fqb;

#define CU_CHECK(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                      \
            return FQB_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

namespace {
constexpr int kAlnCapFast = 8;          // hits kept per read in the fast pass
constexpr int kAlnCapSlow = 1024;       // per read in the overflow pass
constexpr uint32_t kArenaFast = 4096;   // stack entries per lane in the fast pass (bump-allocated; pushes/read ~300 mean); twice that for reads over 128 bases
constexpr uint32_t kArenaMid = 60000;   // overflow tier 1 (still 16-bit bucket heads)
constexpr int kMidBlocks = 16;
// Batch sets (everything the align stage of a batch writes).  With two, the align stage of batch n+2 waits for the later stages of
// batch n; with three it does not, so two align stages can be on the device beside the later stages of a third batch.  On the
// alignment-bound workloads that changes nothing (the device is saturated: 24.5 ms per step either way), but where the align stage
// is one long dependent chain -- filter-bound input, where a few 5,000-step reads are all the search has -- the chains of
// consecutive batches overlap: wgs_mix 5.2 -> 4.2 ms per step (profiles/r02_pipeline_timeline.md).  13 GB per set at 100 bases.
constexpr int kSets = 3;
constexpr int kSpillCap = 16384;        // reads per batch that may need the overflow tiers (rows of d_aln_big); more is a limit error
constexpr int kPenaltyCap = 1 << 16;    // entries of the pairing penalty table (index = insert size <= high_bayesian)
// BatchSet::d_ctrs: [0] n_work [1] queue cursor [2] overflow reads of the fast pass; tier t (1, 2): [4t] n_work [4t+1] cursor
// [4t+2] reads that overflowed tier t; [11] != 0: more than kSpillCap reads overflowed the fast pass
constexpr int kCtrSpillFlag = 11;
// d_status: words of the later stages of the current batch
constexpr int kStSeErr = 0, kStBig = 1 /* [1] n_big [2] over kPairBigMax */, kStSw = 3 /* [3] n_sw [4] cursor */, kStDpErr = 5, kStTuples = 6;
}  // namespace

struct fqb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    HostIndex hidx;
    fqb_gap_opt_t gopt;
    fqb_pe_opt_t popt;
    // device index
    uint4 *d_blocks[2] = {nullptr, nullptr};
    uint32_t *d_sa[2] = {nullptr, nullptr};
    uint8_t *d_pac = nullptr, *d_roll = nullptr;
    int kmer_origin = 0;      // 0 no tables (kmer_thresh == 0), 1 built on the device, 2 uploaded from memory, 3 streamed from <prefix>.rollhash
    DevBwt dbwt[2];
    int32_t *d_maxdiff = nullptr;
    int32_t h_maxdiff[FQB_MAX_READ_LEN + 1];
    // batch buffers.  Everything the align stage of a batch writes exists twice (BatchSet): batch n+1 is prepared and
    // searched on the align stream while batch n goes through pairing / mate rescue / refinement / statistics on the
    // main stream.  The plain members below (bv, wv, d_aln ...) are the view of the CURRENT set (use_set).
    int cap_reads = 0, lpad = 0, stride_cap = 0;
    int n_reads = 0, stride = 0;
    struct BatchSet {
        uint8_t *d_in[4] = {nullptr, nullptr, nullptr, nullptr};   // staging of the host input: bases1, quals1, bases2, quals2
        int32_t *d_lens_in[2] = {nullptr, nullptr};
        BatchView bv; WidthView wv;
        Hit *d_aln = nullptr, *d_aln_big = nullptr;
        int32_t *d_naln = nullptr, *d_spill_slot = nullptr;
        uint32_t *d_overflow = nullptr, *d_work_sorted = nullptr, *d_order_bins = nullptr;
        uint32_t *d_ctrs = nullptr;                 // 16 words, see kCtr*
        int n_reads = 0, stride = 0; bool single_end = false;
        int state = 0;                              // 0 free, 1 loaded, 2 align stage enqueued, 3 being paired (current)
        cudaEvent_t ev_in = nullptr;                // upload complete
        cudaEvent_t ev_free = nullptr;              // prep_kernel has consumed the staging arrays
        cudaEvent_t ev_align = nullptr;             // align stage complete
        cudaEvent_t ev_done = nullptr;              // the later stages no longer read this set
        cudaEvent_t ev_rq[2] = {nullptr, nullptr};  // around the rank-query kernels (width + order + search)
        bool rq_pending = false;
        bool pre_valid = false;                     // staging holds a batch uploaded by fqb_prefetch_pairs, not consumed yet
        const void *pre_key[4] = {nullptr, nullptr, nullptr, nullptr}; int pre_pairs = 0, pre_stride = 0;
    } sets[kSets];
    int cur = 0;                                    // set the stage-level calls work on
    int fifo[kSets] = {}; int n_fifo = 0;         // sets submitted (fqb_submit_pairs) and not collected yet, oldest first
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t align_stream[kSets] = {};   // one per batch set: the search of batch n+1 fills the SMs the draining tail of batch n leaves idle
    cudaEvent_t ev_rows[3] = {nullptr, nullptr, nullptr};   // rows final on the main stream / split done / copies done
    double rq_ms = 0.0; uint64_t rq_launches = 0;   // accumulated device time of the rank-query launches
    BatchView bv;
    WidthView wv;
    Hit *d_aln = nullptr;
    int32_t *d_naln = nullptr;
    uint32_t *d_overflow = nullptr;
    uint32_t *d_ctrs = nullptr;
    uint32_t *d_work_sorted = nullptr, *d_order_bins = nullptr;   // search queue, longest first
    unsigned long long *d_counters = nullptr;
    // search infrastructure
    int n_blocks16 = 0;
    uint4 *d_arena[kSets] = {};       // stack arenas of the fast pass and of the two overflow tiers, per batch set
    uint4 *d_arena_big[kSets] = {};   // (two align stages may be on the device at once)
    uint4 *d_arena_mid[kSets] = {};
    uint32_t arena_fast_alloc[kSets] = {};        // entries per lane d_arena[k] was allocated with
    Hit *d_aln_big = nullptr;
    int32_t *d_spill_slot = nullptr;     // per read: row in d_aln_big or -1
    SearchOpt sopt;
    bool batch_ready = false;
    uint64_t n_launches = 0;
    // Per-batch parameters that depend on infer_isize live on the DEVICE (BatchCtl): the pair stage copies the insert-size
    // histogram to pinned memory, a host callback in the stream (cudaLaunchHostFunc) runs the libm arithmetic and writes
    // the parameters back, and the kernels that follow read them through pointers -- no host synchronisation per batch.
    struct BatchCtl {
        uint64_t rng_calls;               // draws of the drand48 stream consumed by earlier batches (libbwa/bwase.c:33-36)
        uint64_t pad_;
        fqb_isize_t last_ii;              // infer_isize's fallback (src/BwtMapper.cpp:780-781)
        fqb_isize_t cur_ii;
        PairParams pp;
        SwParams sw;
    };
    BatchCtl *d_ctl = nullptr, *h_ctl = nullptr;          // h_ctl: pinned master copy (written by the callback)
    struct PairXfer { uint64_t totals[2]; uint32_t hist[kIsizeBins + 1]; } *h_xfer = nullptr;   // pinned, device -> callback
    int32_t *h_penalty = nullptr;                          // pinned, kPenaltyCap entries
    uint32_t *d_status = nullptr, *h_status = nullptr;     // 16 status words of the batch in the later stages, see kSt*
    uint32_t *h_ctrs = nullptr;                            // pinned copy of the current set's counters (overflow tiers)
    bool status_pending = false;
    std::string cb_error;                                  // set by the host callback (read after a synchronisation)
    uint64_t tuples_bound = 0;                             // upper bound of the pile-up entries on the device
    // FQB_TIMELINE=1 (development): timing events per batch -- [0] align stage enqueued point reached, [1] search done,
    // [2] later stages start, [3] later stages done -- printed relative to the first by fqb_rows_wait
    std::vector<cudaEvent_t> tl_ev; cudaEvent_t tl_base = nullptr;
    // multi-GPU (row e): rank in the sharded run, the NCCL communicator of the end-of-run merge, and the hand-off ring's
    // mailboxes (my own, and the next rank's mapped into this address space: IPC or peer access)
    struct RingBox { unsigned long long words[7]; unsigned int seq, pad_; };
    int comm_rank = 0, comm_world = 1;
    ncclComm_t nccl = nullptr;
    RingBox *ring_inbox = nullptr, *ring_next = nullptr; bool ring_next_ipc = false;
    uint32_t ring_epoch = 0;                  // bumped by fqb_reset_stream (once per file, on every rank): a mailbox never matches a sequence number of an earlier file
    uint64_t prefetch_hits = 0;                            // batches fqb_stage_load found already uploaded by fqb_prefetch_pairs
    // paired-end resolution stage
    fqb_read_t *d_rows = nullptr, *d_rows_split = nullptr;
    PeScratch pesc = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t *d_hist = nullptr;          // kIsizeBins + 1 (last = max_len)
    int32_t *d_penalty = nullptr; size_t penalty_cap = 0;
    int32_t *d_log_n = nullptr;
    uint32_t *d_big_list = nullptr; uint64_t *d_pair_scratch = nullptr;
    std::vector<uint32_t> h_hist;
    uint64_t rng_x0 = 0, rng_calls = 0;  // srand48(bns->seed) stream position (src/BwtMapper.cpp:1817)
    fqb_isize_t last_ii, cur_ii;
    bool align_done = false, pair_done = false, dp_done = false;
    bool single_end = false;                     // the resident batch came without second reads (SingleEndMapper)
    uint32_t *d_sw_list = nullptr, *d_refine_list = nullptr;   // each: work list followed by its retry list
    void *d_sw_huge = nullptr;                                 // buffers of the very-wide-window mate-rescue path
    uint32_t *d_dpctr = nullptr;                             // [0..3] SW list/cursors, [4..7] refine list/cursors, [8] error
    DpPool dp_pool = {nullptr, nullptr, 0, 0, 0};
    // statistics rows (a12-a14)
    bool stats_open = false, stats_done = false;
    std::string target_bed;                      // --targetRegion, set before fqb_stats_open
    long long files_closed[6] = {0, 0, 0, 0, 0, 0}; // device totals already attributed to finished files
    bool files_merged = false;                   // the per-file counters hold the sums over all ranks (fqb_comm_merge_stats)
    StatsTables stabs;
    ContigDev *d_ctg = nullptr;
    uint32_t *d_site = nullptr; int32_t *d_marker = nullptr;
    uint32_t *d_depth = nullptr;                 // depth | q20 | q30, n_sites each
    unsigned long long *d_emp = nullptr;         // [4][256] + isize_dist[4096] + scalars[8]
    uint32_t *d_contig_ctr = nullptr;            // [n_ctg][4] + first[n_ctg]
    unsigned long long *d_dup_keys = nullptr; uint32_t dup_cap = 0;
    PileupTuple *d_tuples = nullptr; uint32_t *d_ntuples = nullptr; uint32_t tuple_cap = 0;
    PairStat *d_pstat = nullptr;
    uint64_t pairs_seen = 0;                     // global pair index of the next batch
    std::vector<PileupColumn> pileup;
    std::vector<PileupTuple> tuples_host;          // pile-up entries drained from the device (own batches + imported ones)
    unsigned long long *d_keys_imp = nullptr; size_t cap_keys_imp = 0;      // receive buffer of other ranks' duplicate keys (fqb_comm_merge_stats)
    PileupTuple *d_tuples_imp = nullptr; size_t n_tuples_imp = 0, cap_tuples_imp = 0;   // entries imported from other handles, kept on the device until the files are written
    std::vector<FileCounters> files;
    // BAM emission (row f1)
    BgzfWriter bam; bool bam_open = false; BamContext bam_ctx;
    fqb_handle *bam_owner = nullptr;       // sharded run in one process: records go to this handle's BAM file (fqb_bam_attach)
    MultiOut *d_multi_out = nullptr; uint32_t *d_multi_list = nullptr, *d_multi_ctr = nullptr; uint32_t multi_cap = 0;
    fqb_read_t *h_bam_rows = nullptr; size_t h_bam_rows_cap = 0;
    // the reference's paired reader reuses its per-slot rseq buffers every second batch without clearing them, and SetSamRecord's
    // "no match" branch prints such a buffer over the full read length: [parity of the batch][end][slot * stride]
    std::vector<uint8_t> rseq_shadow[2][2]; size_t rseq_stride = 0; uint64_t bam_batches = 0;
    std::ofstream isize_table;
    std::string isize_table_path;
    std::vector<std::pair<uint64_t, uint64_t>> isize_table_idx;   // (first global pair, bytes) of every emitted batch, in emission order
    fqb_read_t *h_rows = nullptr; PairStat *h_pstat = nullptr; size_t h_rows_cap = 0;     // pinned staging of fqb_stats_emit
    // Host phases (formatting + writing) of the last fqb_stats_emit / fqb_bam_emit when FQB_ASYNC_EMIT is set: they run on
    // their own thread while the caller submits the next batch, and are joined before their staging is written again (drain_post)
    std::future<void> post_stats, post_bam;
    std::mutex post_m; std::string post_err;
    uint64_t h_pstat_first = ~0ull;        // first global pair of the batch h_pstat holds
    uint32_t arena_fast = kArenaFast, arena_mid = kArenaMid;   // FQB_DEBUG_ARENA_FAST/_MID shrink them (tests of the overflow tiers)
};

static void use_set(fqb_handle *h, int si) {
    fqb_handle::BatchSet &B = h->sets[si];
    h->cur = si;
    h->bv = B.bv; h->wv = B.wv; h->d_aln = B.d_aln; h->d_aln_big = B.d_aln_big; h->d_naln = B.d_naln; h->d_spill_slot = B.d_spill_slot;
    h->d_overflow = B.d_overflow; h->d_work_sorted = B.d_work_sorted; h->d_order_bins = B.d_order_bins; h->d_ctrs = B.d_ctrs;
    h->n_reads = B.n_reads; h->stride = B.stride; h->single_end = B.single_end;
}

static void sync_all(fqb_handle *h) {
    for (cudaStream_t a : h->align_stream) if (a) cudaStreamSynchronize(a);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d2h_stream) cudaStreamSynchronize(h->d2h_stream);
}

static void free_batch(fqb_handle *h) {
    sync_all(h);
    for (auto &B : h->sets) {
        for (auto &p : B.d_in) { cudaFree(p); p = nullptr; }
        for (auto &p : B.d_lens_in) { cudaFree(p); p = nullptr; }
        cudaFree(B.bv.codes); cudaFree(B.bv.qual); cudaFree(B.bv.len); cudaFree(B.bv.full_len);
        cudaFree(B.bv.filtered); cudaFree(B.bv.n_ambig); cudaFree(B.bv.work); cudaFree(B.d_work_sorted);
        cudaFree(B.wv.w); cudaFree(B.wv.sw);
        cudaFree(B.d_aln); cudaFree(B.d_aln_big); cudaFree(B.d_naln); cudaFree(B.d_overflow); cudaFree(B.d_spill_slot);
        memset(&B.bv, 0, sizeof(B.bv)); memset(&B.wv, 0, sizeof(B.wv));
        B.d_aln = B.d_aln_big = nullptr; B.d_naln = B.d_spill_slot = nullptr; B.d_overflow = B.d_work_sorted = nullptr;
        B.state = 0; B.pre_valid = false; B.rq_pending = false;
    }
    h->n_fifo = 0;
    cudaFree(h->d_rows); cudaFree(h->d_rows_split); cudaFree(h->pesc.packed); cudaFree(h->pesc.scanned); cudaFree(h->pesc.scan_tmp); cudaFree(h->pesc.cum_extra);
    cudaFree(h->pesc.multi_list); cudaFree(h->d_big_list); cudaFree(h->d_sw_list); cudaFree(h->d_refine_list); cudaFree(h->d_pstat);
    h->d_sw_list = h->d_refine_list = nullptr; h->d_pstat = nullptr;
    h->d_rows = nullptr; h->pesc.packed = h->pesc.scanned = h->pesc.scan_tmp = h->pesc.cum_extra = nullptr; h->pesc.multi_list = nullptr; h->d_big_list = nullptr;
    h->cap_reads = 0;
    use_set(h, 0);
}

static int ensure_batch(fqb_handle *h, int n_reads, int stride) {
    if (n_reads <= h->cap_reads && stride <= h->stride_cap) return FQB_OK;
    free_batch(h);
    int cap = n_reads > 2 * FQB_BATCH_PAIRS ? n_reads : (n_reads > 65536 ? 2 * FQB_BATCH_PAIRS : 65536 * 2);
    if (stride < h->stride_cap) stride = h->stride_cap;
    int lpad = (stride + 15) & ~15;
    for (auto &B : h->sets) {
        for (int i = 0; i < 4; ++i) CU_CHECK(cudaMalloc(&B.d_in[i], (size_t)(cap / 2) * stride));
        for (int i = 0; i < 2; ++i) CU_CHECK(cudaMalloc(&B.d_lens_in[i], (size_t)(cap / 2) * 4));
        CU_CHECK(cudaMalloc(&B.bv.codes, (size_t)cap * lpad));
        CU_CHECK(cudaMalloc(&B.bv.qual, (size_t)cap * lpad));
        CU_CHECK(cudaMalloc(&B.bv.len, (size_t)cap * 4));
        CU_CHECK(cudaMalloc(&B.bv.full_len, (size_t)cap * 4));
        CU_CHECK(cudaMalloc(&B.bv.filtered, (size_t)cap));
        CU_CHECK(cudaMalloc(&B.bv.n_ambig, (size_t)cap));
        CU_CHECK(cudaMalloc(&B.bv.work, (size_t)cap * 4));
        CU_CHECK(cudaMalloc(&B.d_work_sorted, (size_t)cap * 4));
        B.wv.wstride = stride + 1;
        B.wv.sstride = h->gopt.seed_len + 1;
        CU_CHECK(cudaMalloc(&B.wv.w, (size_t)cap * 2 * B.wv.wstride * 4));
        CU_CHECK(cudaMalloc(&B.wv.sw, (size_t)cap * 2 * B.wv.sstride * 4));
        CU_CHECK(cudaMalloc(&B.d_aln, (size_t)cap * kAlnCapFast * sizeof(Hit)));
        CU_CHECK(cudaMalloc(&B.d_aln_big, (size_t)kSpillCap * kAlnCapSlow * sizeof(Hit)));
        CU_CHECK(cudaMalloc(&B.d_naln, (size_t)cap * 4));
        CU_CHECK(cudaMalloc(&B.d_overflow, (size_t)cap * 3 * 4));
        CU_CHECK(cudaMalloc(&B.d_spill_slot, (size_t)cap * 4));
    }
    CU_CHECK(cudaMalloc(&h->d_rows, (size_t)cap * sizeof(fqb_read_t)));
    CU_CHECK(cudaMalloc(&h->d_rows_split, (size_t)cap * sizeof(fqb_read_t)));
    CU_CHECK(cudaMalloc(&h->pesc.packed, (size_t)cap * 8));
    CU_CHECK(cudaMalloc(&h->pesc.scanned, (size_t)cap * 8));
    CU_CHECK(cudaMalloc(&h->pesc.scan_tmp, ((size_t)cap / 1024 + 2) * 2 * 8));
    CU_CHECK(cudaMalloc(&h->pesc.cum_extra, (size_t)cap * 8));
    CU_CHECK(cudaMalloc(&h->pesc.multi_list, (size_t)cap * 4));
    CU_CHECK(cudaMalloc(&h->d_big_list, (size_t)kPairBigMax * 4));
    CU_CHECK(cudaMalloc(&h->d_sw_list, (size_t)cap * 4));        // n_pairs entries + n_pairs retry entries
    CU_CHECK(cudaMalloc(&h->d_refine_list, (size_t)cap * 8));    // n_reads entries + n_reads retry entries
    CU_CHECK(cudaMalloc(&h->d_pstat, (size_t)(cap / 2) * sizeof(PairStat)));
    h->cap_reads = cap; h->lpad = lpad; h->stride_cap = stride;
    use_set(h, 0);
    return FQB_OK;
}

extern "C" {

static int create_common(fqb_handle *h, int device, fqb_handle **out);
static int check_device(int device) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        set_error("no CUDA device: fastquick_b200 has no CPU fallback");
        return FQB_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) { set_error("bad device ordinal"); return FQB_ERR_ARG; }
    return FQB_OK;
}

int fqb_create(const char *index_prefix, const fqb_gap_opt_t *gopt, const fqb_pe_opt_t *popt, int device, fqb_handle **out) {
    if (!index_prefix || !out) { set_error("null argument"); return FQB_ERR_ARG; }
    if (int rc = check_device(device)) return rc;
    fqb_handle *h = new fqb_handle();
    if (gopt) h->gopt = *gopt; else fqb_gap_opt_default(&h->gopt);
    if (popt) h->popt = *popt; else fqb_pe_opt_default(&h->popt);
    std::string err;
    if (!load_index(index_prefix, false, h->hidx, err)) { set_error(err); delete h; return FQB_ERR_IO; }
    return create_common(h, device, out);
}

// same engine over an index built in memory from synthetic flanks (bench/tests; skips the 3 GiB .rollhash round trip)
int fqb_create_from_synth(const fqb_synth *s, const fqb_gap_opt_t *gopt, const fqb_pe_opt_t *popt, int device, fqb_handle **out) {
    if (!s || !out) { set_error("null argument"); return FQB_ERR_ARG; }
    if (int rc = check_device(device)) return rc;
    fqb_handle *h = new fqb_handle();
    if (gopt) h->gopt = *gopt; else fqb_gap_opt_default(&h->gopt);
    if (popt) h->popt = *popt; else fqb_pe_opt_default(&h->popt);
    build_index_from_flanks(fqb_synth_flanks_internal(s), h->gopt.kmer_thresh != 0 && getenv("FQB_ROLLHASH_FROM_FILE"), h->hidx);
    return create_common(h, device, out);
}

static int create_common(fqb_handle *h, int device, fqb_handle **out) {
    h->device = device;
    if (const char *e = getenv("FQB_DEBUG_ARENA_FAST")) h->arena_fast = (uint32_t)atoi(e);
    if (const char *e = getenv("FQB_DEBUG_ARENA_MID")) h->arena_mid = (uint32_t)atoi(e);
    if ((uint64_t)h->hidx.bwt[0].seq_len + 1 >= (1ull << kWidthBits)) {
        set_error("reduced reference too long for the packed width table (limit 2^27 bases)");
        delete h; return FQB_ERR_LIMIT;
    }
    fill_maxdiff_table(h->gopt, h->h_maxdiff);
    for (int l = 0; l <= FQB_MAX_READ_LEN; ++l)
        if (h->h_maxdiff[l] > 30) { set_error("max_diff > 30 is not supported"); delete h; return FQB_ERR_LIMIT; }
    if (h->popt.n_multi < 0 || h->popt.n_multi + 1 > FQB_MAX_MULTI || h->popt.N_multi < 0 || h->popt.N_multi + 1 > FQB_MAX_MULTI) {
        set_error("--n_multi / --N_multi above 10 are not supported (a result row reports at most 11 hits)"); delete h; return FQB_ERR_LIMIT;
    }
    if (h->gopt.max_gapo > 14 || h->gopt.max_gape > 30 || h->gopt.seed_len > 255 || h->gopt.s_mm < 1 || h->gopt.s_gapo < 1 || h->gopt.s_gape < 1) {
        set_error("gap options outside the supported range"); delete h; return FQB_ERR_LIMIT;
    }
#define CU_CHECK_H(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); fqb_destroy(h); return FQB_ERR_CUDA; } } while (0)
    CU_CHECK_H(cudaSetDevice(device));
    {   // the later stages run at a higher priority than the next batch's align stage: their blocks go first when SM room frees up
        // (lo = the numerically largest = least urgent value, which is also what a stream created without a priority gets)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const int main_prio = getenv("FQB_FLAT_PRIORITY") ? lo : hi;
        CU_CHECK_H(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, main_prio));
        CU_CHECK_H(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CU_CHECK_H(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        for (auto &a : h->align_stream) CU_CHECK_H(cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, lo));
    }
    for (auto &e : h->ev_rows) CU_CHECK_H(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &B : h->sets) {
        CU_CHECK_H(cudaEventCreateWithFlags(&B.ev_in, cudaEventDisableTiming));
        CU_CHECK_H(cudaEventCreateWithFlags(&B.ev_free, cudaEventDisableTiming));
        CU_CHECK_H(cudaEventCreateWithFlags(&B.ev_align, cudaEventDisableTiming));
        CU_CHECK_H(cudaEventCreateWithFlags(&B.ev_done, cudaEventDisableTiming));
        CU_CHECK_H(cudaEventCreate(&B.ev_rq[0])); CU_CHECK_H(cudaEventCreate(&B.ev_rq[1]));
        CU_CHECK_H(cudaMalloc(&B.d_ctrs, 16 * 4));
        CU_CHECK_H(cudaMemset(B.d_ctrs, 0, 16 * 4));
        CU_CHECK_H(cudaMalloc(&B.d_order_bins, 32 * 4));
    }
    CU_CHECK_H(cudaMalloc(&h->d_ctl, sizeof(fqb_handle::BatchCtl)));
    CU_CHECK_H(cudaMallocHost(&h->h_ctl, sizeof(fqb_handle::BatchCtl)));
    CU_CHECK_H(cudaMallocHost(&h->h_xfer, sizeof(fqb_handle::PairXfer)));
    CU_CHECK_H(cudaMallocHost(&h->h_penalty, (size_t)kPenaltyCap * 4));
    CU_CHECK_H(cudaMalloc(&h->d_penalty, (size_t)kPenaltyCap * 4));
    CU_CHECK_H(cudaMalloc(&h->d_status, 16 * 4));
    CU_CHECK_H(cudaMemset(h->d_status, 0, 16 * 4));
    CU_CHECK_H(cudaMallocHost(&h->h_status, 16 * 4));
    CU_CHECK_H(cudaMallocHost(&h->h_ctrs, 16 * 4));
    memset(h->h_status, 0, 16 * 4); memset(h->h_ctrs, 0, 16 * 4);
    memset(h->h_ctl, 0, sizeof(fqb_handle::BatchCtl));
    for (int s = 0; s < 2; ++s) {
        std::vector<Block32> blocks;
        relayout_bwt(h->hidx.bwt[s], blocks);
        CU_CHECK_H(cudaMalloc(&h->d_blocks[s], blocks.size() * sizeof(Block32)));
        CU_CHECK_H(cudaMemcpy(h->d_blocks[s], blocks.data(), blocks.size() * sizeof(Block32), cudaMemcpyHostToDevice));
        CU_CHECK_H(cudaMalloc(&h->d_sa[s], h->hidx.bwt[s].sa.size() * 4));
        CU_CHECK_H(cudaMemcpy(h->d_sa[s], h->hidx.bwt[s].sa.data(), h->hidx.bwt[s].sa.size() * 4, cudaMemcpyHostToDevice));
        DevBwt &d = h->dbwt[s];
        d.blocks = h->d_blocks[s]; d.sa = h->d_sa[s];
        d.primary = h->hidx.bwt[s].primary; d.seq_len = h->hidx.bwt[s].seq_len;
        for (int i = 0; i < 5; ++i) d.L2[i] = h->hidx.bwt[s].L2[i];
        d.n_blocks = (uint32_t)blocks.size();
    }
    CU_CHECK_H(cudaMalloc(&h->d_pac, h->hidx.pac.size()));
    CU_CHECK_H(cudaMemcpy(h->d_pac, h->hidx.pac.data(), h->hidx.pac.size(), cudaMemcpyHostToDevice));
    CU_CHECK_H(cudaMalloc(&h->d_maxdiff, sizeof(h->h_maxdiff)));
    CU_CHECK_H(cudaMemcpy(h->d_maxdiff, h->h_maxdiff, sizeof(h->h_maxdiff), cudaMemcpyHostToDevice));
    CU_CHECK_H(cudaMalloc(&h->pesc.totals, 4 * 8));
    h->pesc.err_flag = h->d_status + kStSeErr;
    CU_CHECK_H(cudaMalloc(&h->d_hist, (kIsizeBins + 1) * 4));
    {
        int32_t g[256];
        fill_log_n(g);
        CU_CHECK_H(cudaMalloc(&h->d_log_n, sizeof(g)));
        CU_CHECK_H(cudaMemcpy(h->d_log_n, g, sizeof(g), cudaMemcpyHostToDevice));
    }
    h->rng_x0 = lcg_seed(h->hidx.seed); h->rng_calls = 0;
    h->last_ii.avg = h->last_ii.std = -1.0; h->last_ii.ap_prior = 0; h->last_ii.low = h->last_ii.high = h->last_ii.high_bayesian = h->last_ii.pad_ = 0;
    h->cur_ii = h->last_ii;
    h->h_ctl->rng_calls = 0; h->h_ctl->last_ii = h->last_ii; h->h_ctl->cur_ii = h->last_ii;
    CU_CHECK_H(cudaMemcpy(h->d_ctl, h->h_ctl, sizeof(fqb_handle::BatchCtl), cudaMemcpyHostToDevice));
    CU_CHECK_H(cudaMalloc(&h->d_dpctr, 12 * 4));
    CU_CHECK_H(cudaMalloc(&h->d_counters, 512 * 8));          // 16 counters + the timeline slots of the development build (make kstats)
    CU_CHECK_H(cudaMemset(h->d_counters, 0, 512 * 8));
    if (h->gopt.kmer_thresh != 0) {     // 6 x 512 MiB membership tables (BwtIndexer::roll_hash_table)
        const size_t total = kRollTableBytes * kNumRollTables;
        CU_CHECK_H(cudaMalloc(&h->d_roll, total));
        KmerBuildInputs kin;
        if (h->hidx.rollhash.size() == total) {
            CU_CHECK_H(cudaMemcpy(h->d_roll, h->hidx.rollhash.data(), total, cudaMemcpyHostToDevice));
            std::vector<uint8_t>().swap(h->hidx.rollhash);
            h->kmer_origin = 2;
        } else if (!getenv("FQB_ROLLHASH_FROM_FILE") && kmer_build_inputs(h->hidx, kin)) {
            // built here from the flank text (row f4): same bits as <prefix>.rollhash, without the 3 GiB read
            const size_t nj = kin.special.size();
            std::vector<int64_t> jf(nj), jl(nj); std::vector<int32_t> jn(nj), jt(nj);
            for (size_t j = 0; j < nj; ++j) { jf[j] = kin.special[j].first; jl[j] = kin.special[j].last; jn[j] = kin.special[j].len; jt[j] = kin.special[j].table; }
            uint8_t *d_codes = nullptr, *d_alleles = nullptr, *d_sp = nullptr; int64_t *d_off = nullptr, *d_jf = nullptr, *d_jl = nullptr; int32_t *d_jn = nullptr, *d_jt = nullptr;
            auto up = [&](auto **dst, const void *src, size_t bytes) -> cudaError_t {
                cudaError_t e = cudaMalloc((void **)dst, bytes + 64);
                return e != cudaSuccess ? e : cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
            };
            CU_CHECK_H(up(&d_codes, kin.codes.data(), kin.codes.size()));
            CU_CHECK_H(up(&d_alleles, kin.alleles.data(), kin.alleles.size()));
            CU_CHECK_H(up(&d_off, kin.offsets.data(), kin.offsets.size() * 8));
            CU_CHECK_H(up(&d_sp, kin.special_codes.data(), kin.special_codes.size()));
            CU_CHECK_H(up(&d_jf, jf.data(), nj * 8)); CU_CHECK_H(up(&d_jl, jl.data(), nj * 8));
            CU_CHECK_H(up(&d_jn, jn.data(), nj * 4)); CU_CHECK_H(up(&d_jt, jt.data(), nj * 4));
            CU_CHECK_H(cudaMemsetAsync(h->d_roll, 0, total, h->stream));
            KmerBuildView kv;
            kv.codes = d_codes; kv.offset = d_off; kv.alleles = d_alleles; kv.n_flanks = (int)h->hidx.contigs.size();
            kv.n_bases = h->hidx.l_pac; kv.tables = reinterpret_cast<uint32_t *>(h->d_roll);
            launch_kmer_build(kv, h->stream);
            KmerSpecialView sv;
            sv.codes = d_sp; sv.first = d_jf; sv.last = d_jl; sv.len = d_jn; sv.table = d_jt; sv.n_jobs = (int)nj; sv.tables = kv.tables;
            launch_kmer_build_special(sv, h->stream);
            CU_CHECK_H(cudaGetLastError());
            CU_CHECK_H(cudaStreamSynchronize(h->stream));
            cudaFree(d_codes); cudaFree(d_alleles); cudaFree(d_off); cudaFree(d_sp); cudaFree(d_jf); cudaFree(d_jl); cudaFree(d_jn); cudaFree(d_jt);
            h->kmer_origin = 1;
        } else {                        // streamed from disk (BwtIndexer::ReadRollHashTable)
            FILE *fp = fopen(h->hidx.rollhash_path.c_str(), "rb");
            if (!fp) { set_error("cannot open " + h->hidx.rollhash_path); fqb_destroy(h); return FQB_ERR_IO; }
            const size_t chunk = 64u << 20;
            std::vector<uint8_t> buf(chunk);
            size_t done = 0;
            while (done < total) {
                size_t want = total - done < chunk ? total - done : chunk;
                if (fread(buf.data(), 1, want, fp) != want) { fclose(fp); set_error("short .rollhash"); fqb_destroy(h); return FQB_ERR_IO; }
                CU_CHECK_H(cudaMemcpy(h->d_roll + done, buf.data(), want, cudaMemcpyHostToDevice));
                done += want;
            }
            fclose(fp);
            h->kmer_origin = 3;
        }
    }
    *out = h;
    return FQB_OK;
}

static int drain_post(fqb_handle *h, bool stats, bool bam);
static void comm_release(fqb_handle *h);

int fqb_kmer_tables_origin(const fqb_handle *h) { return h ? h->kmer_origin : 0; }

int fqb_kmer_tables_fetch(fqb_handle *h, uint64_t offset, uint64_t n_bytes, uint8_t *out) {
    if (!h || !out) { set_error("null argument"); return FQB_ERR_ARG; }
    if (!h->d_roll) { set_error("no k-mer tables (kmer_thresh == 0)"); return FQB_ERR_ARG; }
    if (offset + n_bytes > kRollTableBytes * kNumRollTables) { set_error("range outside the k-mer tables"); return FQB_ERR_ARG; }
    cudaSetDevice(h->device);
    cudaError_t e = cudaMemcpy(out, h->d_roll + offset, n_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error(std::string("cudaMemcpy: ") + cudaGetErrorString(e)); return FQB_ERR_CUDA; }
    return FQB_OK;
}

void fqb_destroy(fqb_handle *h) {
    if (h) drain_post(h, true, true);
    if (!h) return;
    cudaSetDevice(h->device);
    free_batch(h);
    for (int s = 0; s < 2; ++s) { cudaFree(h->d_blocks[s]); cudaFree(h->d_sa[s]); }
    cudaFree(h->d_pac); cudaFree(h->d_roll); cudaFree(h->d_maxdiff); cudaFree(h->d_counters);
    for (int k = 0; k < kSets; ++k) { cudaFree(h->d_arena[k]); cudaFree(h->d_arena_mid[k]); cudaFree(h->d_arena_big[k]); }
    for (auto &B : h->sets) {
        cudaFree(B.d_ctrs); cudaFree(B.d_order_bins);
        for (cudaEvent_t e : {B.ev_in, B.ev_free, B.ev_align, B.ev_done, B.ev_rq[0], B.ev_rq[1]}) if (e) cudaEventDestroy(e);
    }
    comm_release(h);
    cudaFree(h->d_ctl); cudaFree(h->d_status); cudaFreeHost(h->h_ctl); cudaFreeHost(h->h_xfer); cudaFreeHost(h->h_penalty); cudaFreeHost(h->h_status); cudaFreeHost(h->h_ctrs);
    // statistics accumulators (fqb_stats_open)
    cudaFree(h->d_ctg); cudaFree(h->d_site); cudaFree(h->d_marker); cudaFree(h->d_depth); cudaFree(h->d_emp); cudaFree(h->d_contig_ctr);
    cudaFree(h->d_dup_keys); cudaFree(h->d_tuples); cudaFree(h->d_ntuples);
    cudaFree(h->pesc.totals); cudaFree(h->d_hist); cudaFree(h->d_penalty); cudaFree(h->d_log_n); cudaFree(h->d_pair_scratch); cudaFree(h->dp_pool.ints); cudaFree(h->dp_pool.bytes); cudaFree(h->d_dpctr); cudaFree(h->d_sw_huge);
    cudaFreeHost(h->h_rows); cudaFreeHost(h->h_pstat); cudaFreeHost(h->h_bam_rows);
    cudaFree(h->d_multi_out); cudaFree(h->d_multi_list); cudaFree(h->d_multi_ctr); cudaFree(h->d_tuples_imp); cudaFree(h->d_keys_imp);
    if (h->bam_open) { std::string e; h->bam.close(e); }
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
    for (cudaStream_t a : h->align_stream) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    for (auto &e : h->ev_rows) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int fqb_index_info(const fqb_handle *h, int64_t *l_pac, int32_t *n_contigs, uint32_t *primary, uint32_t *seed) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    if (l_pac) *l_pac = h->hidx.l_pac;
    if (n_contigs) *n_contigs = (int32_t)h->hidx.contigs.size();
    if (primary) { primary[0] = h->hidx.bwt[0].primary; primary[1] = h->hidx.bwt[1].primary; }
    if (seed) *seed = h->hidx.seed;
    return FQB_OK;
}

// ---- the per-batch chain ------------------------------------------------------
// Every stage below only ENQUEUES work (kernels, copies, one host callback) on a stream; nothing waits for the device.
// Errors a stage can only detect on the device are written to status words and read back at the end of the chain;
// check_status() turns them into return codes after the caller-visible synchronisation point.
}  // extern "C" (reopened below)

// overflow tier t (1 or 2) takes over the reads the previous pass gave up on: its counters, and (tier 1) the rows of the
// big hit buffer the reads will write to
static __global__ void tier_setup_kernel(uint32_t *ctrs, int tier, const uint32_t *list, int32_t *spill_slot, uint32_t spill_cap) {
    uint32_t n = ctrs[4 * (tier - 1) + 2];
    if (tier == 1 && n > spill_cap) { n = spill_cap; if (blockIdx.x == 0 && threadIdx.x == 0) ctrs[kCtrSpillFlag] = 1; }
    if (tier == 1)
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) spill_slot[list[j]] = (int32_t)j;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrs[4 * tier] = n; ctrs[4 * tier + 1] = 0; ctrs[4 * tier + 2] = 0; }
}

static void harvest_rq(fqb_handle *h, fqb_handle::BatchSet &B, bool wait) {
    if (!B.rq_pending) return;
    if (wait) cudaEventSynchronize(B.ev_rq[1]);
    else if (cudaEventQuery(B.ev_rq[1]) != cudaSuccess) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, B.ev_rq[0], B.ev_rq[1]) == cudaSuccess) { h->rq_ms += ms; ++h->rq_launches; }
    B.rq_pending = false;
}

// a1..a5 of the batch in set si: prep -> widths -> queue order -> search -> overflow tiers (device-driven: the tier kernels
// always run and find their work lists empty in the common case)
static int enqueue_align(fqb_handle *h, int si, cudaStream_t st) {
    fqb_handle::BatchSet &B = h->sets[si];
    harvest_rq(h, B, true);
    CU_CHECK(cudaMemsetAsync(B.d_ctrs, 0, 16 * 4, st));
    PrepParams pp;
    pp.trim_qual = h->gopt.trim_qual; pp.kmer_thresh = h->gopt.kmer_thresh; pp.is_il13 = h->gopt.is_il13; pp.roll = h->d_roll;
    launch_prep(B.bv, pp, st);
    CU_CHECK(cudaEventRecord(B.ev_free, st));        // the staging arrays may be refilled from here on
    h->n_launches += 3;
    CU_CHECK(cudaEventRecord(B.ev_rq[0], st));
    launch_width(B.bv, B.wv, h->dbwt, h->gopt.seed_len, B.bv.work, B.bv.n_work, B.n_reads, h->d_counters, st);

    h->sopt = make_search_opt(h->gopt, B.stride);
    if (h->sopt.n_buckets > 128) { set_error("more than 128 score buckets"); return FQB_ERR_LIMIT; }
    // Fast-pass arena: 4,096 entries per lane hold every read of a 2 x 100 bp batch; of 150-base reads (twice the pushes) about one
    // per batch outgrows that, and its overflow tier is a single 10-ms chain behind the search (2x150_10k: 54.8 ms per step
    // against 51.4 with 8,192 entries, 52.1 with 16,384).  FQB_DEBUG_ARENA_FAST pins the size (tests of the tiers).
    const bool arena_pinned = getenv("FQB_DEBUG_ARENA_FAST") != nullptr;
    const uint32_t want_fast = arena_pinned ? h->arena_fast : (B.stride > 128 ? 2 * kArenaFast : kArenaFast);
    if (h->d_arena[si] && h->arena_fast_alloc[si] < want_fast) { CU_CHECK(cudaFree(h->d_arena[si])); h->d_arena[si] = nullptr; }   // longer reads than before
    if (!h->d_arena[si]) {
        if (!h->n_blocks16) h->n_blocks16 = search_grid_blocks(h->sopt.n_buckets, true, h->device);
        CU_CHECK(cudaMalloc(&h->d_arena[si], (size_t)h->n_blocks16 * kSearchThreads * want_fast * sizeof(uint4)));
        h->arena_fast_alloc[si] = want_fast;
        if (!h->d_arena_mid[si]) CU_CHECK(cudaMalloc(&h->d_arena_mid[si], (size_t)kMidBlocks * kSearchThreads * h->arena_mid * sizeof(uint4)));
        if (!h->d_arena_big[si]) CU_CHECK(cudaMalloc(&h->d_arena_big[si], (size_t)kSearchThreads * ((size_t)h->gopt.max_entries + 64) * sizeof(uint4)));
    }
    SearchParams sp;
    sp.bwt[0] = h->dbwt[0]; sp.bwt[1] = h->dbwt[1];
    sp.opt = h->sopt; sp.maxdiff = h->d_maxdiff; sp.seed_len_opt = h->gopt.seed_len;
    sp.work = B.bv.work; sp.n_work = B.d_ctrs; sp.cursor = B.d_ctrs + 1;
    sp.pops_out = nullptr;
    if (!getenv("FQB_NO_ORDER")) {
        launch_order(B.bv, B.wv, B.bv.work, B.bv.n_work, B.n_reads, B.d_order_bins, B.d_work_sorted, st);
        h->n_launches += 2;
        sp.work = B.d_work_sorted;
    }
    sp.arena = h->d_arena[si]; sp.arena_cap = h->arena_fast_alloc[si];
    sp.aln = B.d_aln; sp.aln_cap = kAlnCapFast; sp.n_aln = B.d_naln; sp.aln_row = nullptr;
    sp.overflow = B.d_overflow; sp.n_overflow = B.d_ctrs + 2;
    sp.counters = h->d_counters;
    CU_CHECK(cudaMemsetAsync(B.d_naln, 0, (size_t)B.n_reads * 4, st));
    CU_CHECK(cudaMemsetAsync(B.d_spill_slot, 0xff, (size_t)B.n_reads * 4, st));
    launch_search(B.bv, B.wv, sp, true, false, h->n_blocks16, st);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaEventRecord(B.ev_rq[1], st));
    B.rq_pending = true;
    // Rare reads whose stack or hit list outgrew the fast pass are redone from scratch with deeper arenas:
    // tier 1 = 60k entries per lane on 16 blocks, tier 2 = as deep as the reference allows (max_entries), one block.
    for (int tier = 1; tier <= 2; ++tier) {
        const uint32_t *list = tier == 1 ? B.d_overflow : B.d_overflow + h->cap_reads;
        uint32_t *d_c = B.d_ctrs + 4 * tier;
        tier_setup_kernel<<<32, 256, 0, st>>>(B.d_ctrs, tier, list, B.d_spill_slot, (uint32_t)kSpillCap);
        // widths were mutated by gap_shadow before the overflow: recompute them for these reads
        launch_width(B.bv, B.wv, h->dbwt, h->gopt.seed_len, list, d_c, kSpillCap, nullptr, st);
        SearchParams s2 = sp;
        s2.work = list; s2.n_work = d_c; s2.cursor = d_c + 1;
        s2.arena = tier == 1 ? h->d_arena_mid[si] : h->d_arena_big[si];
        s2.arena_cap = tier == 1 ? h->arena_mid : (uint32_t)h->gopt.max_entries + 64;
        s2.aln = B.d_aln_big; s2.aln_cap = kAlnCapSlow; s2.aln_row = B.d_spill_slot;
        s2.overflow = B.d_overflow + h->cap_reads * (tier == 1 ? 1 : 2); s2.n_overflow = d_c + 2;
        s2.counters = nullptr; s2.pops_out = nullptr;
        launch_search(B.bv, B.wv, s2, tier == 1, true, tier == 1 ? kMidBlocks : 1, st);
        h->n_launches += 3;
    }
    CU_CHECK(cudaGetLastError());
    B.state = 2;
    return FQB_OK;
}

// infer_isize + the fallbacks of src/BwtMapper.cpp:779-786 and everything derived from the estimate, on the host with
// glibc's libm, in stream order: runs on a CUDA callback thread between the histogram's copy to pinned
// memory and the copy of the parameters back to the device.  Must not call the CUDA API.
static void CUDART_CB pair_host_callback(void *user) {
    fqb_handle *h = static_cast<fqb_handle *>(user);
    fqb_handle::BatchCtl &C = *h->h_ctl;
    C.rng_calls += h->h_xfer->totals[0];
    fqb_isize_t ii;
    infer_isize_hist(h->h_xfer->hist, (int)h->h_xfer->hist[kIsizeBins], h->popt.ap_prior, (int64_t)h->hidx.bwt[0].seq_len, ii);
    if (ii.avg < 0.0 && C.last_ii.avg > 0.0) ii = C.last_ii;
    if (h->popt.force_isize) { ii.low = ii.high = 0; ii.avg = ii.std = -1.0; }
    C.cur_ii = ii; C.last_ii = ii;
    std::vector<int32_t> pen;
    fill_isize_penalty(ii, pen);
    if (pen.size() > (size_t)kPenaltyCap) { h->cb_error = "insert-size estimate too wide for the pairing penalty table (high_bayesian >= 65536)"; pen.resize(kPenaltyCap); }
    if (!pen.empty()) memcpy(h->h_penalty, pen.data(), pen.size() * 4);
    PairParams &pp = C.pp;
    pp.high = ii.high; pp.high_bayesian = ii.high_bayesian; pp.max_isize = h->popt.max_isize; pp.s_mm = h->gopt.s_mm;
    pp.max_occ = h->popt.max_occ; pp.n_multi = h->popt.n_multi; pp.N_multi = h->popt.N_multi;
    pp.penalty = h->d_penalty; pp.g_log_n = h->d_log_n;
    pp.sw_on = 0;       // mate-rescue candidates are collected by the mate-rescue stage itself
    SwParams &sw = C.sw;
    sw.on = h->popt.is_sw && ii.avg >= 0.0 ? 1 : 0;
    sw.avg = ii.avg; sw.std = ii.std; sw.l_pac = h->hidx.l_pac;
    sw.s_old_add = sw.on ? -4.343 * std::log(ii.ap_prior / h->hidx.l_pac) : 0.0;               // libbwa/bwape.c:577
    sw.s_new_add = (int)(-4.343 * std::log(.5 * std::erfc(M_SQRT1_2 * 1.5) + .499));           // libbwa/bwape.c:578
    h->rng_calls = C.rng_calls; h->last_ii = C.last_ii; h->cur_ii = C.cur_ii;
}
// SingleEndMapper has no insert sizes: only the stream position moves on
static void CUDART_CB se_host_callback(void *user) {
    fqb_handle *h = static_cast<fqb_handle *>(user);
    h->h_ctl->rng_calls += h->h_xfer->totals[0];
    h->rng_calls = h->h_ctl->rng_calls;
}

// the 56 bytes at the head of BatchCtl (rng_calls, pad, last_ii) into the next rank's mailbox, then its sequence number
static __global__ void ring_send_kernel(const unsigned long long *state, volatile unsigned long long *peer_words, volatile unsigned int *peer_seq, unsigned int seq) {
    for (int i = 0; i < 7; ++i) peer_words[i] = state[i];
    __threadfence_system();
    *peer_seq = seq;
}
static __global__ void ring_recv_kernel(unsigned long long *state, const volatile unsigned long long *words, const volatile unsigned int *my_seq, unsigned int seq) {
    while (*my_seq != seq) __nanosleep(100);
    __threadfence_system();
    for (int i = 0; i < 7; ++i) state[i] = words[i];
}
// a6-a9 on the aligned batch (the current set): bwa_cal_pac_pos_pe (src/BwtMapper.cpp:721-907).
// Sharded runs: recv_seq != 0 -> the stream position and last_ii come from this rank's mailbox once it shows recv_seq (the
// wait sits behind the kernels that do not need them); send_seq != 0 -> they go to the next rank's mailbox as soon as the
// host callback has produced them, before the pairing kernels.
static int enqueue_pair(fqb_handle *h, cudaStream_t st, unsigned int recv_seq = 0, unsigned int send_seq = 0) {
    PeView v;
    v.n_reads = h->n_reads;
    v.aln = h->d_aln; v.aln_cap = kAlnCapFast; v.aln_big = h->d_aln_big; v.aln_big_cap = kAlnCapSlow;
    v.spill_slot = h->d_spill_slot; v.n_aln = h->d_naln; v.filtered = h->bv.filtered;
    v.len = h->bv.len; v.full_len = h->bv.full_len; v.rows = h->d_rows; v.single_end = h->single_end ? 1 : 0;
    SeParams sp;
    sp.bwt[0] = h->dbwt[0]; sp.bwt[1] = h->dbwt[1]; sp.maxdiff = h->d_maxdiff; sp.g_log_n = h->d_log_n;
    const RngState rng{h->rng_x0, 0};                  // the stream position is read on the device (d_ctl->rng_calls)
    CU_CHECK(cudaMemsetAsync(h->d_status, 0, 15 * 4, st));      // word 15 is the sticky error word (status_fold_kernel)
    unsigned long long *state = reinterpret_cast<unsigned long long *>(h->d_ctl);
    auto ring_in = [&]() -> int {
        if (!recv_seq) return FQB_OK;
        ring_recv_kernel<<<1, 1, 0, st>>>(state, h->ring_inbox->words, &h->ring_inbox->seq, recv_seq);
        CU_CHECK(cudaMemcpyAsync(h->h_ctl, h->d_ctl, 56, cudaMemcpyDeviceToHost, st));      // the callback works on the pinned master copy
        ++h->n_launches;
        return FQB_OK;
    };
    auto ring_out = [&]() {
        if (!send_seq) return;
        ring_send_kernel<<<1, 1, 0, st>>>(state, h->ring_next->words, &h->ring_next->seq, send_seq);
        ++h->n_launches;
    };
    if (h->single_end) {
        // SingleEndMapper (src/BwtMapper.cpp:1335-1348): bwa_aln2seq_core(..., 1, N_OCC) + bwa_cal_pac_pos; no insert size, no pairing
        launch_se_prepare(v, h->pesc, st);
        if (int rc = ring_in()) return rc;
        launch_se_finish(v, sp, rng, &h->d_ctl->rng_calls, h->pesc, st);
        h->n_launches += 8;
        CU_CHECK(cudaMemcpyAsync(h->h_xfer->totals, h->pesc.totals, 16, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaLaunchHostFunc(st, se_host_callback, h));
        CU_CHECK(cudaMemcpyAsync(&h->d_ctl->rng_calls, &h->h_ctl->rng_calls, 8, cudaMemcpyHostToDevice, st));
        ring_out();
        CU_CHECK(cudaGetLastError());
        return FQB_OK;
    }
    if (h->popt.type != 1) { set_error("only BWA_PET_STD pairing is supported (SOLiD is dead code in the reference)"); return FQB_ERR_ARG; }
    CU_CHECK(cudaMemsetAsync(h->d_hist, 0, (kIsizeBins + 1) * 4, st));
    launch_se_prepare(v, h->pesc, st);
    if (int rc = ring_in()) return rc;
    launch_se_finish(v, sp, rng, &h->d_ctl->rng_calls, h->pesc, st);
    launch_isize_hist(v, h->d_hist, h->d_hist + kIsizeBins, st);
    h->n_launches += 9;
    CU_CHECK(cudaMemcpyAsync(h->h_xfer->hist, h->d_hist, (kIsizeBins + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(h->h_xfer->totals, h->pesc.totals, 16, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaLaunchHostFunc(st, pair_host_callback, h));
    CU_CHECK(cudaMemcpyAsync(h->d_ctl, h->h_ctl, sizeof(fqb_handle::BatchCtl), cudaMemcpyHostToDevice, st));
    ring_out();
    CU_CHECK(cudaMemcpyAsync(h->d_penalty, h->h_penalty, (size_t)kPenaltyCap * 4, cudaMemcpyHostToDevice, st));
    if (!h->d_pair_scratch) CU_CHECK(cudaMalloc(&h->d_pair_scratch, (size_t)kPairBigMax * 8192 * 8));
    launch_pair(v, h->dbwt, &h->d_ctl->pp, h->d_big_list, h->d_status + kStBig, h->d_sw_list, h->d_status + kStSw, st);
    // pairs with many hit positions (repeats): global-memory scratch, one thread per pair (the count stays on the device)
    launch_pair_big(v, h->dbwt, &h->d_ctl->pp, h->d_big_list, h->d_status + kStBig, h->d_pair_scratch, 8192, h->d_sw_list, h->d_status + kStSw, st);
    h->n_launches += 2;
    CU_CHECK(cudaGetLastError());
    return FQB_OK;
}

// a10 + a11 on the paired batch: bwa_paired_sw (libbwa/bwape.c:463) then bwa_refine_gapped (libbwa/bwase.c:339)
static int enqueue_sw_refine(fqb_handle *h, cudaStream_t st) {
    if (!h->dp_pool.ints) {
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
        h->dp_pool.n_blocks = n_sm * 2;
        h->dp_pool.ints_per_lane = 6 * 1025;                 // rows of a <= 1024-column window (mate-rescue windows are ~6 sigma + 2L)
        h->dp_pool.bytes_per_lane = 96 * 1024;               // trace-back + ops
        const size_t lanes = (size_t)h->dp_pool.n_blocks * kDpThreads;
        CU_CHECK(cudaMalloc(&h->dp_pool.ints, lanes * h->dp_pool.ints_per_lane * 4));
        CU_CHECK(cudaMalloc(&h->dp_pool.bytes, lanes * h->dp_pool.bytes_per_lane));
    }
    DpView v;
    v.n_reads = h->n_reads; v.lpad = h->lpad; v.codes = h->bv.codes; v.pac = h->d_pac; v.l_pac = h->hidx.l_pac; v.rows = h->d_rows;
    CU_CHECK(cudaMemsetAsync(h->d_dpctr, 0, 12 * 4, st));
    if (h->popt.is_sw && !h->single_end) {       // whether the batch has an insert-size estimate is known on the device only (d_ctl->sw.on)
        if (!h->d_sw_huge) CU_CHECK(cudaMalloc(&h->d_sw_huge, launch_sw_huge_bytes()));
        launch_sw(v, &h->d_ctl->sw, h->dp_pool, h->d_sw_list, h->d_sw_list + h->cap_reads / 2, h->d_dpctr, h->d_status + kStDpErr, h->d_sw_huge, h->stride, st);
        h->n_launches += 6;      // classify, warp kernel, per-lane retry, and the three kernels of the very-wide-window path
    }
    launch_refine(v, h->dp_pool, h->d_refine_list, h->d_refine_list + h->cap_reads, h->d_dpctr + 4, h->d_status + kStDpErr, h->stride, st);
    h->n_launches += 4;
    CU_CHECK(cudaGetLastError());
    return FQB_OK;
}

// end of a batch's chain: fold what the device flagged for this batch into a sticky word (several chains may complete before
// the host looks) and bring it, the status words and the set's overflow counters to pinned memory
enum : uint32_t { kErrShortRead = 1, kErrSpillCap = 2, kErrDeepOverflow = 4, kErrDrandZero = 8, kErrBigPairs = 16, kErrDp = 32 };
static __global__ void status_fold_kernel(const uint32_t *S, const uint32_t *K, uint32_t *sticky) {
    uint32_t e = 0;
    if (K[kPrepShortFlag]) e |= kErrShortRead;
    if (K[kCtrSpillFlag]) e |= kErrSpillCap;
    if (K[4 * 2 + 2]) e |= kErrDeepOverflow;
    if (S[kStSeErr]) e |= kErrDrandZero;
    if (S[kStBig + 1]) e |= kErrBigPairs;
    if (S[kStDpErr]) e |= kErrDp;
    if (e) atomicOr(sticky, e);
}
static int enqueue_status(fqb_handle *h, cudaStream_t st) {
    status_fold_kernel<<<1, 1, 0, st>>>(h->d_status, h->d_ctrs, h->d_status + 15);
    CU_CHECK(cudaMemcpyAsync(h->h_status, h->d_status, 16 * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(h->h_ctrs, h->d_ctrs, 16 * 4, cudaMemcpyDeviceToHost, st));
    h->status_pending = true;
    return FQB_OK;
}
// after a synchronisation of the main stream: what the device (or the callback) reported for the batches completed since
static int check_status(fqb_handle *h) {
    if (!h->status_pending) return FQB_OK;
    h->status_pending = false;
    if (!h->cb_error.empty()) { set_error(h->cb_error); h->cb_error.clear(); return FQB_ERR_LIMIT; }
    const uint32_t e = h->h_status[15];
    if (!e) return FQB_OK;
    cudaMemsetAsync(h->d_status + 15, 0, 4, h->stream);          // reported: start afresh
    h->h_status[15] = 0;
    if (e & kErrShortRead) set_error("a read is shorter than 96 bases: the reference's k-mer filter reads bases 0..95 whatever the read length and sees stale buffer bytes there (src/BwtIndexer.cpp:443-450); unsupported -- disable the filter (kmer_thresh = 0) for such input");
    else if (e & kErrSpillCap) set_error("more than 16,384 reads of one batch outgrew the fast search pass");
    else if (e & kErrDeepOverflow) set_error("a read overflowed even the max_entries-deep arena or 1024 hits");
    else if (e & kErrDrandZero) set_error("drand48 returned exactly 0.0 for a read with one best interval (draw number fqb_drand48_zero_index(seed) of the file's stream, once per 2^48 draws): unsupported");
    else if (e & kErrBigPairs) set_error("too many repeat-heavy pairs in one batch");
    else set_error("an alignment needed more DP scratch or CIGAR operations than provisioned");
    return FQB_ERR_LIMIT;
}
static int sync_and_check(fqb_handle *h) {
    CU_CHECK(cudaStreamSynchronize(h->stream));
    CU_CHECK(cudaGetLastError());
    return check_status(h);
}

extern "C" {
// ---- stage-level entry points ------------------------------------------------
// upload (or adopt the device pointers of) one batch into set si
// packed_stride = 0: bases are ASCII rows of `stride` bytes; else 2-bit rows of packed_stride bytes (fqb_pack_reads) and
// the quality rows carry the not-ACGT flag in bit 7
static int load_set(fqb_handle *h, int si, cudaStream_t st, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                    const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device, int32_t packed_stride = 0) {
    fqb_handle::BatchSet &B = h->sets[si];
    B.n_reads = 2 * n_pairs; B.stride = stride;
    B.single_end = bases2 == nullptr;
    const uint8_t *src[4] = {bases1, quals1, bases2, quals2};
    const int32_t *lsrc[2] = {lens1, lens2};
    BatchView &b = B.bv;
    b.n_reads = B.n_reads; b.stride_in = stride; b.packed_stride = packed_stride; b.lpad = h->lpad;
    const size_t bytes = (size_t)n_pairs * stride, bytes_bases = (size_t)n_pairs * (packed_stride ? packed_stride : stride);
    if (on_device) {
        b.bases_in[0] = bases1; b.quals_in[0] = quals1; b.bases_in[1] = bases2; b.quals_in[1] = quals2;
        b.lens_in[0] = lens1; b.lens_in[1] = lens2;
    } else {
        if (!packed_stride && B.pre_valid && B.pre_pairs == n_pairs && B.pre_stride == stride && B.pre_key[0] == bases1 && B.pre_key[1] == quals1 &&
            B.pre_key[2] == bases2 && B.pre_key[3] == quals2) {
            CU_CHECK(cudaStreamWaitEvent(st, B.ev_in, 0));          // uploaded ahead of time by fqb_prefetch_pairs
            ++h->prefetch_hits;
        } else {
            CU_CHECK(cudaStreamWaitEvent(st, B.ev_free, 0));
            for (int i = 0; i < 4; ++i) if (src[i]) CU_CHECK(cudaMemcpyAsync(B.d_in[i], src[i], (i & 1) ? bytes : bytes_bases, cudaMemcpyHostToDevice, st));
            for (int i = 0; i < 2; ++i)
                if (lsrc[i]) CU_CHECK(cudaMemcpyAsync(B.d_lens_in[i], lsrc[i], (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        }
        B.pre_valid = false;
        b.bases_in[0] = B.d_in[0]; b.quals_in[0] = B.d_in[1];
        b.bases_in[1] = bases2 ? B.d_in[2] : nullptr; b.quals_in[1] = bases2 ? B.d_in[3] : nullptr;
        b.lens_in[0] = lens1 ? B.d_lens_in[0] : nullptr; b.lens_in[1] = lens2 ? B.d_lens_in[1] : nullptr;
    }
    b.n_work = B.d_ctrs;
    B.state = 1;
    return FQB_OK;
}
static int check_shape(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1, const uint8_t *bases2, const uint8_t *quals2,
                       int32_t packed_stride = 0) {
    if (!h || n_pairs < 0 || stride < 1 || stride > FQB_MAX_READ_LEN) { set_error("bad batch shape"); return FQB_ERR_ARG; }
    if (packed_stride && ((packed_stride & 15) || packed_stride * 4 < stride || packed_stride > FQB_MAX_READ_LEN / 4)) {
        set_error("packed rows are a multiple of 16 bytes that holds `stride` bases (fqb_packed_stride)"); return FQB_ERR_ARG;
    }
    if (!bases1 || !quals1 || (bases2 && !quals2)) { set_error("bases and qualities are required"); return FQB_ERR_ARG; }
    return FQB_OK;
}

static int stage_load_impl(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                           const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device, int32_t packed_stride) {
    if (int rc = check_shape(h, n_pairs, stride, bases1, quals1, bases2, quals2, packed_stride)) return rc;
    if (h->n_fifo) { set_error("fqb_stage_load: batches submitted with fqb_submit_pairs are still in flight (collect them first)"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    int rc = ensure_batch(h, 2 * n_pairs, stride);
    if (rc) return rc;
    // the set holding this batch's prefetched upload, else the current one (unless ITS staging holds another prefetched batch)
    int si = h->cur;
    for (int k = 0; k < kSets; ++k) {
        const fqb_handle::BatchSet &B = h->sets[k];
        if (!on_device && B.pre_valid && B.pre_pairs == n_pairs && B.pre_stride == stride && B.pre_key[0] == bases1 && B.pre_key[1] == quals1 &&
            B.pre_key[2] == bases2 && B.pre_key[3] == quals2) si = k;
    }
    if (!on_device && h->sets[si].pre_valid && h->sets[si].pre_key[0] != bases1) si = (si + 1) % kSets;
    CU_CHECK(cudaStreamWaitEvent(h->stream, h->sets[si].ev_done, 0));
    rc = load_set(h, si, h->stream, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, on_device, packed_stride);
    if (rc) return rc;
    use_set(h, si);
    h->batch_ready = true; h->align_done = h->pair_done = h->dp_done = h->stats_done = false;
    return FQB_OK;
}
int fqb_stage_load(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                   const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device) {
    return stage_load_impl(h, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, on_device, 0);
}
int fqb_stage_load_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride, const uint8_t *packed1, const uint8_t *quals1,
                          const int32_t *lens1, const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2, int on_device) {
    if (packed_stride <= 0) { set_error("fqb_stage_load_packed: packed_stride must be positive"); return FQB_ERR_ARG; }
    return stage_load_impl(h, n_pairs, stride, packed1, quals1, lens1, packed2, quals2, lens2, on_device, packed_stride);
}

// a1..a5 on the resident batch: prep -> widths -> search (+ overflow tiers)
int fqb_stage_align(fqb_handle *h) {
    if (!h || !h->batch_ready) { set_error("no batch loaded"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    int rc = enqueue_align(h, h->cur, h->stream);
    if (rc) return rc;
    if ((rc = enqueue_status(h, h->stream))) return rc;
    if ((rc = sync_and_check(h))) return rc;
    harvest_rq(h, h->sets[h->cur], true);
    h->align_done = true;
    return FQB_OK;
}

// a6-a9 on the aligned batch: bwa_cal_pac_pos_pe (src/BwtMapper.cpp:721-907)
int fqb_stage_pair(fqb_handle *h) {
    if (!h || !h->batch_ready || !h->align_done) { set_error("fqb_stage_pair: run fqb_stage_align first"); return FQB_ERR_STATE; }
    if (h->n_reads > 1024 * 1024) { set_error("fqb_stage_pair: at most 524,288 pairs per batch"); return FQB_ERR_LIMIT; }
    CU_CHECK(cudaSetDevice(h->device));
    int rc = enqueue_pair(h, h->stream);
    if (rc) return rc;
    if ((rc = enqueue_status(h, h->stream))) return rc;
    if ((rc = sync_and_check(h))) return rc;
    h->pair_done = true; h->dp_done = false;
    return FQB_OK;
}

// a10 + a11 on the paired batch: bwa_paired_sw (libbwa/bwape.c:463) then bwa_refine_gapped (libbwa/bwase.c:339)
int fqb_stage_sw_refine(fqb_handle *h) {
    if (!h || !h->pair_done) { set_error("fqb_stage_sw_refine: run fqb_stage_pair first"); return FQB_ERR_STATE; }
    if (h->dp_done) return FQB_OK;
    CU_CHECK(cudaSetDevice(h->device));
    int rc = enqueue_sw_refine(h, h->stream);
    if (rc) return rc;
    if ((rc = enqueue_status(h, h->stream))) return rc;
    if ((rc = sync_and_check(h))) return rc;
    h->dp_done = true;
    return FQB_OK;
}

// ---- statistics rows -------------------------------------------------------------------------
// d_emp: 4 x 256 quality/cycle histograms, InsertSizeDist[4096], 9 scalars (NumPCRDup, NumPairReads, out-of-range inserts,
// TotalFiltered, BwaUnmapped, TotalMAPQ, TotalRetained, NumBase, NumRead), then the number of distinct duplicate keys
static __global__ void add_scalar_kernel(unsigned long long *p, unsigned long long v) { *p += v; }
constexpr int kEmpScalars = 9, kEmpWords = 4 * 256 + 4096 + kEmpScalars;

static int drain_tuples(fqb_handle *h) {
    uint32_t n = 0;
    CU_CHECK(cudaMemcpy(&n, h->d_ntuples, 4, cudaMemcpyDeviceToHost));
    if (n > h->tuple_cap) { set_error("pile-up tuple buffer overflow"); return FQB_ERR_LIMIT; }
    if (!n) return FQB_OK;
    const size_t at = h->tuples_host.size();
    h->tuples_host.resize(at + n);
    CU_CHECK(cudaMemcpy(h->tuples_host.data() + at, h->d_tuples, (size_t)n * sizeof(PileupTuple), cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemset(h->d_ntuples, 0, 4));
    return FQB_OK;
}
// UpdateInfoVecAtMarker's appends in arrival order = (global pair, end, offset on the read), whichever handle saw the pair
static void build_pileup(fqb_handle *h) {
    std::vector<PileupTuple> &t = h->tuples_host;
    if (h->n_tuples_imp) {
        const size_t at = t.size();
        t.resize(at + h->n_tuples_imp);
        cudaMemcpy(t.data() + at, h->d_tuples_imp, h->n_tuples_imp * sizeof(PileupTuple), cudaMemcpyDeviceToHost);
        h->n_tuples_imp = 0;
    }
    std::sort(t.begin(), t.end(), [](const PileupTuple &a, const PileupTuple &b) {
        if (a.key_hi != b.key_hi) return a.key_hi < b.key_hi;
        return a.key_lo < b.key_lo;
    });
    h->pileup.assign(h->stabs.markers.size(), PileupColumn());
    for (const PileupTuple &x : t) {
        PileupColumn &c = h->pileup[x.marker];
        c.seq.push_back("ACGTN"[x.base > 4 ? 4 : x.base]);
        c.qual.push_back((char)x.qual);
        c.cycle.push_back(x.cycle);
        c.maq.push_back(x.mapq);
        c.strand.push_back(x.strand != 0);
    }
}

// StatCollector::SetTargetRegion (src/BwtMapper.cpp:227-228): restrict the regular-site statistics to a BED file
int fqb_stats_set_target_region(fqb_handle *h, const char *bed_path) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    if (h->stats_open) { set_error("fqb_stats_set_target_region: call it before fqb_stats_open"); return FQB_ERR_STATE; }
    h->target_bed = bed_path ? bed_path : "";
    return FQB_OK;
}

// RestoreVcfSites + SetGenomeSize (src/BwtMapper.cpp:225-226): side tables and accumulators
int fqb_stats_open(fqb_handle *h, const char *index_prefix) {
    if (!h || !index_prefix) { set_error("null argument"); return FQB_ERR_ARG; }
    if (h->stats_open) { set_error("fqb_stats_open: statistics are already open on this handle"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    std::string err;
    if (!build_stats_tables(h->hidx, index_prefix, h->gopt, h->target_bed, h->stabs, err)) { set_error(err); return FQB_ERR_IO; }
    const StatsTables &T = h->stabs;
    const size_t nc = T.contigs.size(), ns = T.n_sites ? T.n_sites : 1;
    CU_CHECK(cudaMalloc(&h->d_ctg, nc * sizeof(ContigDev)));
    CU_CHECK(cudaMemcpy(h->d_ctg, T.contigs.data(), nc * sizeof(ContigDev), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&h->d_site, T.site.size() * 4));
    CU_CHECK(cudaMemcpy(h->d_site, T.site.data(), T.site.size() * 4, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&h->d_marker, T.marker_at.size() * 4));
    CU_CHECK(cudaMemcpy(h->d_marker, T.marker_at.data(), T.marker_at.size() * 4, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&h->d_depth, ns * 3 * 4));
    CU_CHECK(cudaMemset(h->d_depth, 0, ns * 3 * 4));
    CU_CHECK(cudaMalloc(&h->d_emp, (kEmpWords + 1) * 8));
    CU_CHECK(cudaMemset(h->d_emp, 0, (kEmpWords + 1) * 8));
    CU_CHECK(cudaMalloc(&h->d_contig_ctr, nc * 5 * 4));
    CU_CHECK(cudaMemset(h->d_contig_ctr, 0, nc * 4 * 4));
    CU_CHECK(cudaMemset(h->d_contig_ctr + nc * 4, 0xff, nc * 4));
    h->dup_cap = 1u << 28;                       // 268M-slot open-addressing set of (start, end) keys (2 GiB): room for the 200M-pair configuration
    CU_CHECK(cudaMalloc(&h->d_dup_keys, (size_t)h->dup_cap * 8));
    CU_CHECK(cudaMemset(h->d_dup_keys, 0, (size_t)h->dup_cap * 8));
    h->tuple_cap = 1u << 25;
    CU_CHECK(cudaMalloc(&h->d_tuples, (size_t)h->tuple_cap * sizeof(PileupTuple)));
    CU_CHECK(cudaMalloc(&h->d_ntuples, 4));
    CU_CHECK(cudaMemset(h->d_ntuples, 0, 4));
    h->pileup.assign(T.markers.size(), PileupColumn());
    h->files.clear();
    for (auto &x : h->files_closed) x = 0;
    h->pairs_seen = 0;
    h->stats_open = true;
    return FQB_OK;
}

// back to the state right after fqb_stats_open: every accumulator zero, no pile-up entries, no duplicate keys (a new run on
// the same handle; bench.py separates its warm-up from its timed region with it)
int fqb_stats_reset(fqb_handle *h) {
    if (!h || !h->stats_open) { set_error("fqb_stats_reset: statistics are not open"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    if (int rc = drain_post(h, true, true)) return rc;
    cudaStream_t st = h->stream;
    const size_t nc = h->stabs.contigs.size(), ns = h->stabs.n_sites ? h->stabs.n_sites : 1;
    CU_CHECK(cudaMemsetAsync(h->d_depth, 0, ns * 3 * 4, st));
    CU_CHECK(cudaMemsetAsync(h->d_emp, 0, (kEmpWords + 1) * 8, st));
    CU_CHECK(cudaMemsetAsync(h->d_contig_ctr, 0, nc * 4 * 4, st));
    CU_CHECK(cudaMemsetAsync(h->d_contig_ctr + nc * 4, 0xff, nc * 4, st));
    CU_CHECK(cudaMemsetAsync(h->d_dup_keys, 0, (size_t)h->dup_cap * 8, st));
    CU_CHECK(cudaMemsetAsync(h->d_ntuples, 0, 4, st));
    CU_CHECK(cudaStreamSynchronize(st));
    h->tuples_host.clear(); h->n_tuples_imp = 0; h->tuples_bound = 0;
    h->pairs_seen = 0;
    for (auto &x : h->files_closed) x = 0;
    h->files.clear(); h->files_merged = false;
    return FQB_OK;
}

// The FileStatCollector counters live on the device as running totals; a file's own counters are the totals at its
// end minus the totals when it began (collector.AddFSC(FSC), src/BwtMapper.cpp:254).
// joins the deferred host phases; reports the first failure one of them met
static int drain_post(fqb_handle *h, bool stats, bool bam) {
    if (stats && h->post_stats.valid()) h->post_stats.get();
    if (bam && h->post_bam.valid()) h->post_bam.get();
    std::lock_guard<std::mutex> l(h->post_m);
    if (!h->post_err.empty()) { set_error(h->post_err); h->post_err.clear(); return FQB_ERR_IO; }
    return FQB_OK;
}
// Host phases run inline unless FQB_ASYNC_EMIT is set: on the 16-core GPU box the threaded variant lost to the inline one
// (text + BAM 7.1e5 against 9.5e5 pairs/s, BGZF + BAM 4.4e5 against 9.4e5: the formatter threads take the cores the feeder's
// inflate / parse workers need), so it stays an option for hosts with cores to spare.
static bool emit_inline() { static const bool v = getenv("FQB_ASYNC_EMIT") == nullptr; return v; }

static int close_current_file(fqb_handle *h) {
    if (int rc = drain_post(h, true, true)) return rc;
    if (h->files.empty() || h->files_merged) return FQB_OK;       // after fqb_comm_merge_stats the per-file counters are final
    CU_CHECK(cudaStreamSynchronize(h->stream));
    unsigned long long sc[kEmpScalars];
    CU_CHECK(cudaMemcpy(sc, h->d_emp + 4 * 256 + 4096, sizeof sc, cudaMemcpyDeviceToHost));
    FileCounters &F = h->files.back();
    long long *dst[6] = {&F.TotalFiltered, &F.BwaUnmapped, &F.TotalMAPQ, &F.TotalRetained, &F.NumBase, &F.NumRead};
    for (int k = 0; k < 6; ++k) { *dst[k] = (long long)sc[3 + k] - h->files_closed[k]; h->files_closed[k] = (long long)sc[3 + k]; }
    return FQB_OK;
}

// the current file's FileStatCollector counters so far (src/BwtMapper.cpp:2116-2122 prints them when a file is done):
// out6 = TotalFiltered, BwaUnmapped (both in pairs; the reference prints them x 2), TotalMAPQ, TotalRetained, NumBase, NumRead
int fqb_stats_file_counters(fqb_handle *h, int64_t *out6) {
    if (!h || !h->stats_open || !out6) { set_error("fqb_stats_file_counters: statistics are not open"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    unsigned long long sc[kEmpScalars];
    CU_CHECK(cudaMemcpy(sc, h->d_emp + 4 * 256 + 4096, sizeof sc, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; ++k) out6[k] = (int64_t)sc[3 + k] - h->files_closed[k];
    return FQB_OK;
}

// a new FASTQ pair: FileStatCollector FSC(fq1, fq2) (src/BwtMapper.cpp:249-254); the first call (re)creates <out_prefix>.InsertSizeTable
int fqb_stats_begin_file(fqb_handle *h, const char *out_prefix, const char *fq1, const char *fq2) {
    if (!h || !h->stats_open) { set_error("fqb_stats_begin_file: call fqb_stats_open first"); return FQB_ERR_STATE; }
    { int rc = close_current_file(h); if (rc) return rc; }
    if (out_prefix && !h->isize_table.is_open()) {
        h->isize_table_path = std::string(out_prefix) + ".InsertSizeTable";
        h->isize_table.open(h->isize_table_path);
        h->isize_table_idx.clear();
        if (!h->isize_table) { set_error("cannot write the InsertSizeTable"); return FQB_ERR_IO; }
    }
    for (auto &a : h->rseq_shadow) for (auto &b : a) std::fill(b.begin(), b.end(), 0);      // PairEndMapper allocates fresh read buffers per file
    h->bam_batches = 0;
    FileCounters f;
    f.FileName1 = fq1 ? fq1 : ""; f.FileName2 = fq2 ? fq2 : "";
    h->files.push_back(f);
    return fqb_reset_stream(h);
}

// a12 + a13 on the finished batch: AddAlignment's decision tree per pair, then the per-base accumulation
static int enqueue_stats(fqb_handle *h, cudaStream_t st) {
    if (h->files.empty()) h->files.push_back(FileCounters());
    const size_t np = (size_t)h->n_reads / 2, nc = h->stabs.contigs.size(), ns = h->stabs.n_sites ? h->stabs.n_sites : 1;
    if (h->pairs_seen + np > (1ull << 31)) { set_error("more than 2^31 read pairs through one handle: the arrival-order keys would wrap"); return FQB_ERR_LIMIT; }
    // pile-up entries accumulate on the device; a batch appends at most one per aligned base of a marker site, far fewer than
    // 2 per read.  Drain them to the host (the one synchronisation of this stage, every few dozen batches) before they can overflow.
    if (h->tuples_bound + (uint64_t)h->n_reads * 2 > h->tuple_cap) {
        CU_CHECK(cudaStreamSynchronize(st));
        int rc = drain_tuples(h);
        if (rc) return rc;
        h->tuples_bound = 0;
    }
    h->tuples_bound += (uint64_t)h->n_reads * 2;
    StatsView v;
    v.n_reads = h->n_reads; v.lpad = h->lpad; v.codes = h->bv.codes; v.qual = h->bv.qual; v.rows = h->d_rows; v.pstat = h->d_pstat;
    v.ctg = h->d_ctg; v.n_ctg = (int)nc; v.pair_base = (uint32_t)h->pairs_seen; v.cal_dup = 1; v.pac = h->d_pac;
    StatAccum A;
    A.contig_ctr = h->d_contig_ctr; A.contig_first = h->d_contig_ctr + nc * 4;
    A.isize_dist = h->d_emp + 4 * 256; A.scalars = h->d_emp + 4 * 256 + 4096;
    A.dup_keys = h->d_dup_keys; A.dup_cap = h->dup_cap; A.dup_count = h->d_emp + kEmpWords;
    BaseTables B;
    B.site = h->d_site; B.marker = h->d_marker; B.depth = h->d_depth; B.q20 = h->d_depth + ns; B.q30 = h->d_depth + 2 * ns;
    B.emp = h->d_emp; B.tuples = h->d_tuples; B.n_tuples = h->d_ntuples; B.tuple_cap = h->tuple_cap;
    launch_classify(v, A, st);
    launch_bases(v, B, st);
    h->n_launches += 2;
    CU_CHECK(cudaGetLastError());
    const unsigned long long n_in = h->single_end ? np : 2ull * np;        // reads that came from the FASTQ file(s)
    h->files.back().NumRead += (long long)n_in;
    add_scalar_kernel<<<1, 1, 0, st>>>(h->d_emp + 4 * 256 + 4096 + 8, n_in);     // NumRead, kept on the device too so that sharded runs sum it
    h->pairs_seen += np;
    return FQB_OK;
}
int fqb_stage_stats(fqb_handle *h) {
    if (!h || !h->stats_open || !h->dp_done) { set_error("fqb_stage_stats: needs fqb_stats_open and a batch through fqb_stage_sw_refine"); return FQB_ERR_STATE; }
    if (h->stats_done) return FQB_OK;
    CU_CHECK(cudaSetDevice(h->device));
    int rc = enqueue_stats(h, h->stream);
    if (rc) return rc;
    h->stats_done = true;
    return FQB_OK;
}

// text emission for the last batch: one InsertSizeTable line per retained pair (ProcessPairStatus's fout lines).
// names: n_pairs rows of name_stride bytes (NUL-padded), or null for the synthetic r%011lld names.
int fqb_stats_emit2(fqb_handle *h, const char *names, const char *names2, int32_t name_stride) {
    if (!names2) names2 = names;
    if (!h || !h->stats_done) { set_error("fqb_stats_emit: run fqb_stage_stats first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    if (int rc = drain_post(h, true, true)) return rc;                  // the previous batch's host phases read the staging written below
    const size_t np = (size_t)h->n_reads / 2;
    if (np > h->h_rows_cap) {
        cudaFreeHost(h->h_rows); cudaFreeHost(h->h_pstat);
        h->h_rows = nullptr; h->h_pstat = nullptr; h->h_rows_cap = 0;
        CU_CHECK(cudaMallocHost(&h->h_rows, 2 * np * sizeof(fqb_read_t)));
        CU_CHECK(cudaMallocHost(&h->h_pstat, np * sizeof(PairStat)));
        h->h_rows_cap = np;
    }
    const uint64_t first = h->pairs_seen - np;
    CU_CHECK(cudaMemcpyAsync(h->h_rows, h->d_rows, 2 * np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(cudaMemcpyAsync(h->h_pstat, h->d_pstat, np * sizeof(PairStat), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    h->h_pstat_first = first;
    if (!h->isize_table.is_open()) return FQB_OK;
    // host phase: format on several host threads (contiguous slices of the batch), write the slices in order.  `names` must
    // stay valid until the next fqb_stats_emit / fqb_bam_emit / fqb_stats_finish / fqb_bam_close call on this handle returns.
    auto host_phase = [h, np, names, names2, name_stride, first]() {
        unsigned nthr = std::thread::hardware_concurrency();
        if (nthr < 1) nthr = 1;
        if (nthr > 16) nthr = 16;
        if (np < 4096) nthr = 1;
        std::vector<std::string> parts(nthr);
        auto work = [&](unsigned t) {
            std::string &o = parts[t];
            const size_t lo = np * t / nthr, hi = np * (t + 1) / nthr;
            o.reserve((hi - lo) * 96);
            char buf[64];
            std::string nm;
            for (size_t i = lo; i < hi; ++i) {
                const PairStat &ps = h->h_pstat[i];
                if (ps.line_kind == 0) continue;
                const char *name;
                // the line of a pair whose first read is unmapped carries the SECOND read's name (q->name, src/StatCollector.cpp:695,708)
                const char *src = ps.line_kind == 2 ? names2 : names;
                if (src) { nm.assign(src + i * (size_t)name_stride, strnlen(src + i * (size_t)name_stride, (size_t)name_stride)); name = nm.c_str(); }
                else { snprintf(buf, sizeof buf, "r%011llu", (unsigned long long)(first + i)); name = buf; }
                append_isize_line(h->stabs, ps, h->h_rows[2 * i], h->h_rows[2 * i + 1], name, o);
            }
        };
        if (nthr == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthr; ++t) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        uint64_t bytes = 0;
        for (auto &o : parts) { h->isize_table.write(o.data(), (std::streamsize)o.size()); bytes += o.size(); }
        h->isize_table_idx.emplace_back(first, bytes);
        if (!h->isize_table) { std::lock_guard<std::mutex> l(h->post_m); if (h->post_err.empty()) h->post_err = "write error on the InsertSizeTable"; }
    };
    if (emit_inline()) { host_phase(); return drain_post(h, true, true); }
    h->post_stats = std::async(std::launch::async, host_phase);
    return FQB_OK;
}

int fqb_stats_emit(fqb_handle *h, const char *names, int32_t name_stride) { return fqb_stats_emit2(h, names, nullptr, name_stride); }

int fqb_emit_sync(fqb_handle *h) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    return drain_post(h, true, true);
}

// Sharded runs: every rank writes the InsertSizeTable lines of its own batches.  close_table() finishes a rank's file
// and leaves "<file>.idx" (first global pair and byte count of each batch) next to it; merge_tables(), called on the
// rank that will run fqb_stats_finish, splices all ranks' batches back into file order, so the table (and the
// InsertSizeEstimator that reads it) is exactly that of an unsharded run.
int fqb_stats_close_table(fqb_handle *h) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    if (int rc = drain_post(h, true, true)) return rc;
    if (h->isize_table.is_open()) h->isize_table.close();
    if (h->isize_table_path.empty()) return FQB_OK;
    std::ofstream idx(h->isize_table_path + ".idx");
    for (auto &e : h->isize_table_idx) idx << e.first << "\t" << e.second << "\n";
    if (!idx) { set_error("cannot write the InsertSizeTable index"); return FQB_ERR_IO; }
    return FQB_OK;
}

int fqb_stats_merge_tables(fqb_handle *h, const char *const *other_prefixes, int32_t n_others) {
    if (!h || n_others < 0 || (n_others && !other_prefixes)) { set_error("bad arguments"); return FQB_ERR_ARG; }
    int rc = fqb_stats_close_table(h);
    if (rc) return rc;
    if (h->isize_table_path.empty()) { set_error("fqb_stats_merge_tables: no InsertSizeTable open on this handle"); return FQB_ERR_STATE; }
    std::vector<std::pair<uint64_t, std::string>> chunks;
    auto load = [&](const std::string &path) -> bool {
        std::ifstream idx(path + ".idx"), tab(path, std::ios::binary);
        if (!idx || !tab) return false;
        uint64_t first, bytes;
        while (idx >> first >> bytes) {
            std::string c(bytes, '\0');
            if (bytes && !tab.read(&c[0], (std::streamsize)bytes)) return false;
            chunks.emplace_back(first, std::move(c));
        }
        return true;
    };
    if (!load(h->isize_table_path)) { set_error("cannot read " + h->isize_table_path); return FQB_ERR_IO; }
    for (int i = 0; i < n_others; ++i)
        if (!load(std::string(other_prefixes[i]) + ".InsertSizeTable")) { set_error(std::string("cannot read the InsertSizeTable of ") + other_prefixes[i]); return FQB_ERR_IO; }
    std::stable_sort(chunks.begin(), chunks.end(), [](const std::pair<uint64_t, std::string> &a, const std::pair<uint64_t, std::string> &b) { return a.first < b.first; });
    std::ofstream out(h->isize_table_path, std::ios::binary | std::ios::trunc);
    for (auto &c : chunks) out << c.second;
    if (!out) { set_error("cannot rewrite " + h->isize_table_path); return FQB_ERR_IO; }
    return FQB_OK;
}

// ProcessCore (src/StatCollector.cpp:2012-2028): gathers the accumulators and writes every summary file
int fqb_stats_finish(fqb_handle *h, const char *out_prefix) {
    if (!h || !h->stats_open || !out_prefix) { set_error("fqb_stats_finish: call fqb_stats_open first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    int rc = drain_post(h, true, true);
    if (rc) return rc;
    rc = drain_tuples(h);
    if (rc) return rc;
    if (h->isize_table.is_open()) h->isize_table.close();
    const StatsTables &T = h->stabs;
    const size_t nc = T.contigs.size(), ns = T.n_sites ? T.n_sites : 1;
    StatsTotals S;
    std::vector<uint32_t> d(ns * 3);
    CU_CHECK(cudaMemcpy(d.data(), h->d_depth, ns * 3 * 4, cudaMemcpyDeviceToHost));
    S.depth.assign(d.begin(), d.begin() + ns); S.q20.assign(d.begin() + ns, d.begin() + 2 * ns); S.q30.assign(d.begin() + 2 * ns, d.end());
    std::vector<unsigned long long> e(kEmpWords + 1);
    CU_CHECK(cudaMemcpy(e.data(), h->d_emp, e.size() * 8, cudaMemcpyDeviceToHost));
    S.emp.assign(e.begin(), e.begin() + 1024);
    S.isize_dist.assign(e.begin() + 1024, e.begin() + 1024 + 4096);
    const unsigned long long *sc = e.data() + 1024 + 4096;
    if (sc[2]) { set_error("an insert size fell outside InsertSizeDist[4096] (the reference would write out of bounds), or the duplicate-key table is full"); return FQB_ERR_LIMIT; }
    S.num_pcr_dup = sc[0]; S.num_pair_reads = sc[1];
    std::vector<uint32_t> cc(nc * 5);
    CU_CHECK(cudaMemcpy(cc.data(), h->d_contig_ctr, nc * 5 * 4, cudaMemcpyDeviceToHost));
    S.contig_ctr.assign(cc.begin(), cc.begin() + nc * 4); S.contig_first.assign(cc.begin() + nc * 4, cc.end());
    build_pileup(h);
    S.pileup = h->pileup;
    rc = close_current_file(h);           // the last file's counters (in a sharded run: after the totals were imported)
    if (rc) return rc;
    S.files = h->files;
    std::string err;
    if (!write_summary_files(T, S, h->gopt, out_prefix, err)) { set_error(err); return FQB_ERR_IO; }
    return FQB_OK;
}

// ---- multi-GPU plumbing (row e): reads shard by batch with the index replicated; what crosses GPUs is
// (1) the position of the drand48 stream + last_ii, handed from the rank that owns batch b to the owner of b+1, and
// (2) at the end, the integer accumulators (NCCL reduce through torch.distributed on buffers exported here).
// the state lives on the device (BatchCtl) with a pinned master copy the pair stage's callback updates: both calls first
// wait for the batches in flight
static int push_stream_state(fqb_handle *h) {
    h->h_ctl->rng_calls = h->rng_calls; h->h_ctl->last_ii = h->last_ii;
    CU_CHECK(cudaMemcpyAsync(h->d_ctl, h->h_ctl, offsetof(fqb_handle::BatchCtl, cur_ii), cudaMemcpyHostToDevice, h->stream));
    return FQB_OK;
}
int fqb_get_stream_state(fqb_handle *h, uint64_t *rng_calls, fqb_isize_t *last_ii) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    if (rng_calls) *rng_calls = h->rng_calls;
    if (last_ii) *last_ii = h->last_ii;
    return FQB_OK;
}
int fqb_set_stream_state(fqb_handle *h, uint64_t rng_calls, const fqb_isize_t *last_ii) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    h->rng_calls = rng_calls;
    if (last_ii) h->last_ii = *last_ii;
    return push_stream_state(h);
}
int fqb_set_pair_base(fqb_handle *h, uint64_t first_pair) {     // global index of the next batch's first pair (pile-up / contig order keys)
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    h->pairs_seen = first_pair;
    return FQB_OK;
}

// accumulator groups: 0 = depth|q20|q30 (u32, sum), 1 = quality/cycle histograms + InsertSizeDist + scalars (u64, sum),
// 2 = per-contig counters (u32, sum), 3 = per-contig first-touch order (u32, min)
static int stats_group(fqb_handle *h, int which, void **ptr, size_t *bytes) {
    const size_t nc = h->stabs.contigs.size(), ns = h->stabs.n_sites ? h->stabs.n_sites : 1;
    switch (which) {
    case 0: *ptr = h->d_depth; *bytes = ns * 3 * 4; return FQB_OK;
    case 1: *ptr = h->d_emp; *bytes = (size_t)kEmpWords * 8; return FQB_OK;
    case 2: *ptr = h->d_contig_ctr; *bytes = nc * 4 * 4; return FQB_OK;
    case 3: *ptr = h->d_contig_ctr + nc * 4; *bytes = nc * 4; return FQB_OK;
    default: set_error("bad accumulator group"); return FQB_ERR_ARG;
    }
}
int fqb_stats_group_bytes(fqb_handle *h, int which, uint64_t *bytes) {
    if (!h || !h->stats_open) { set_error("stats not open"); return FQB_ERR_STATE; }
    void *p; size_t b;
    int rc = stats_group(h, which, &p, &b);
    if (rc == FQB_OK) *bytes = b;
    return rc;
}
int fqb_stats_export(fqb_handle *h, int which, void *dst_device) {
    if (!h || !h->stats_open) { set_error("stats not open"); return FQB_ERR_STATE; }
    void *p; size_t b;
    int rc = stats_group(h, which, &p, &b);
    if (rc) return rc;
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaMemcpyAsync(dst_device, p, b, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    return FQB_OK;
}
int fqb_stats_import(fqb_handle *h, int which, const void *src_device) {
    if (!h || !h->stats_open) { set_error("stats not open"); return FQB_ERR_STATE; }
    void *p; size_t b;
    int rc = stats_group(h, which, &p, &b);
    if (rc) return rc;
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaMemcpyAsync(p, src_device, b, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    return FQB_OK;
}

}  // extern "C" (reopened below)
static __global__ void dup_compact_kernel(const unsigned long long *keys, uint32_t cap, unsigned long long *out, unsigned long long *n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && keys[i]) out[atomicAdd(n_out, 1ull)] = keys[i];
}
// insert another handle's distinct keys; a key already present is one more duplicated pair (NumPCRDup += 2)
static __global__ void dup_merge_kernel(unsigned long long *keys, uint32_t cap, const unsigned long long *in, unsigned long long n,
                                        unsigned long long *n_distinct, unsigned long long *num_pcr_dup) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = in[i];
    uint32_t hh = (uint32_t)(hash64(key) % cap);
    for (uint32_t probes = 0; probes < cap; ++probes) {
        const unsigned long long old = atomicCAS(keys + hh, 0ull, key);
        if (old == 0) { atomicAdd(n_distinct, 1ull); return; }
        if (old == key) { atomicAdd(num_pcr_dup, 2ull); return; }
        hh = hh + 1 == cap ? 0 : hh + 1;
    }
    atomicAdd(num_pcr_dup + 2, 1ull);      // scalars[2]: table full, reported as a limit error by fqb_stats_finish
}
// ---- cross-rank duplicate count of a sharded run, partitioned: key -> owner rank (hash), every owner counts the keys it
// receives from more than one rank.  A key held by m ranks is m - 1 more duplicated pairs (NumPCRDup += 2 each).
constexpr int kMaxRanks = 64;
__device__ __forceinline__ uint32_t key_owner(unsigned long long key, int W) { return (uint32_t)((hash64(key) >> 32) % (unsigned long long)W); }
static __global__ void dup_part_count_kernel(const unsigned long long *keys, uint32_t cap, int W, unsigned long long *counts) {
    __shared__ unsigned int sh[kMaxRanks];
    if (threadIdx.x < kMaxRanks) sh[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k) atomicAdd(&sh[key_owner(k, W)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < W && sh[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}
// out: the keys grouped by owner (segment d starts at the sum of counts[0..d)); cursors: W zeroed words
static __global__ void dup_part_scatter_kernel(const unsigned long long *keys, uint32_t cap, int W, const unsigned long long *counts,
                                               unsigned long long *cursors, unsigned long long *out) {
    __shared__ unsigned long long start[kMaxRanks];
    if (threadIdx.x == 0) { unsigned long long acc = 0; for (int d = 0; d < W; ++d) { start[d] = acc; acc += counts[d]; } }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (!k) continue;
        const uint32_t d = key_owner(k, W);
        out[start[d] + atomicAdd(cursors + d, 1ull)] = k;
    }
}
static __global__ void dup_cross_kernel(unsigned long long *table, unsigned long long mask, const unsigned long long *in, unsigned long long n,
                                        unsigned long long *num_pcr_dup) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = in[i];
    unsigned long long hh = hash64(key) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(table + hh, 0ull, key);
        if (old == 0) return;
        if (old == key) { atomicAdd(num_pcr_dup, 2ull); return; }
        hh = (hh + 1) & mask;
    }
}
extern "C" {
// Variable-size statistics state of a sharded run: which = 0 pile-up entries (sizeof(PileupTuple) = 20 bytes each),
// which = 1 the distinct PCR-duplicate keys (8 bytes each).  The handle that writes the files imports the other
// handles' state AFTER the fixed-size groups (fqb_stats_import overwrites the scalar counters the key merge adds to).
int fqb_stats_var_count(fqb_handle *h, int which, uint64_t *n) {
    if (!h || !h->stats_open || !n) { set_error("fqb_stats_var_count: call fqb_stats_open first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    if (which == 0) {
        uint32_t nd = 0;
        CU_CHECK(cudaMemcpy(&nd, h->d_ntuples, 4, cudaMemcpyDeviceToHost));
        if (nd > h->tuple_cap) { set_error("pile-up tuple buffer overflow"); return FQB_ERR_LIMIT; }
        *n = h->tuples_host.size() + nd + h->n_tuples_imp; return FQB_OK;
    }
    if (which == 1) {
        unsigned long long c = 0;
        CU_CHECK(cudaMemcpy(&c, h->d_emp + kEmpWords, 8, cudaMemcpyDeviceToHost));
        *n = c; return FQB_OK;
    }
    set_error("bad variable-size group"); return FQB_ERR_ARG;
}
// dst / src may be host or device memory (unified addressing decides the copy direction)
int fqb_stats_var_export(fqb_handle *h, int which, void *dst, uint64_t cap) {
    uint64_t n = 0;
    int rc = fqb_stats_var_count(h, which, &n);
    if (rc) return rc;
    if (n > cap) { set_error("fqb_stats_var_export: destination too small"); return FQB_ERR_ARG; }
    if (!n) return FQB_OK;
    if (which == 0) {
        const size_t nh = h->tuples_host.size(), ni = h->n_tuples_imp, nd = (size_t)n - nh - ni;
        if (nh) CU_CHECK(cudaMemcpyAsync(dst, h->tuples_host.data(), nh * sizeof(PileupTuple), cudaMemcpyDefault, h->stream));
        if (nd) CU_CHECK(cudaMemcpyAsync(static_cast<char *>(dst) + nh * sizeof(PileupTuple), h->d_tuples, nd * sizeof(PileupTuple), cudaMemcpyDefault, h->stream));
        if (ni) CU_CHECK(cudaMemcpyAsync(static_cast<char *>(dst) + (nh + nd) * sizeof(PileupTuple), h->d_tuples_imp, ni * sizeof(PileupTuple), cudaMemcpyDefault, h->stream));
        CU_CHECK(cudaStreamSynchronize(h->stream));
        return FQB_OK;
    }
    unsigned long long *tmp = nullptr, *cnt = nullptr;
    CU_CHECK(cudaMalloc(&tmp, n * 8)); CU_CHECK(cudaMalloc(&cnt, 8)); CU_CHECK(cudaMemsetAsync(cnt, 0, 8, h->stream));
    dup_compact_kernel<<<(h->dup_cap + 255) / 256, 256, 0, h->stream>>>(h->d_dup_keys, h->dup_cap, tmp, cnt);
    ++h->n_launches;
    cudaError_t e = cudaMemcpyAsync(dst, tmp, n * 8, cudaMemcpyDefault, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp); cudaFree(cnt);
    if (e != cudaSuccess) { set_error(std::string("CUDA: ") + cudaGetErrorString(e)); return FQB_ERR_CUDA; }
    return FQB_OK;
}
int fqb_stats_var_import(fqb_handle *h, int which, const void *src, uint64_t n) {
    if (!h || !h->stats_open || (n && !src)) { set_error("fqb_stats_var_import: call fqb_stats_open first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    if (!n) return FQB_OK;
    if (which == 0) {
        // staged on the device (a device-to-device copy when the source is a gathered NCCL buffer); they reach the host once, in fqb_stats_finish
        if (h->n_tuples_imp + n > h->cap_tuples_imp) {
            const size_t cap = (h->n_tuples_imp + n) * 2;
            PileupTuple *nb = nullptr;
            CU_CHECK(cudaMalloc(&nb, cap * sizeof(PileupTuple)));
            if (h->n_tuples_imp) CU_CHECK(cudaMemcpyAsync(nb, h->d_tuples_imp, h->n_tuples_imp * sizeof(PileupTuple), cudaMemcpyDeviceToDevice, h->stream));
            CU_CHECK(cudaStreamSynchronize(h->stream));
            cudaFree(h->d_tuples_imp);
            h->d_tuples_imp = nb; h->cap_tuples_imp = cap;
        }
        CU_CHECK(cudaMemcpyAsync(h->d_tuples_imp + h->n_tuples_imp, src, n * sizeof(PileupTuple), cudaMemcpyDefault, h->stream));
        CU_CHECK(cudaStreamSynchronize(h->stream));
        h->n_tuples_imp += n;
        return FQB_OK;
    }
    if (which != 1) { set_error("bad variable-size group"); return FQB_ERR_ARG; }
    unsigned long long *tmp = nullptr;
    CU_CHECK(cudaMalloc(&tmp, n * 8));
    cudaError_t e = cudaMemcpyAsync(tmp, src, n * 8, cudaMemcpyDefault, h->stream);
    if (e == cudaSuccess) {
        dup_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_dup_keys, h->dup_cap, tmp, n, h->d_emp + kEmpWords,
                                                                             h->d_emp + (4 * 256 + 4096));
        ++h->n_launches;
        e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error(std::string("CUDA: ") + cudaGetErrorString(e)); return FQB_ERR_CUDA; }
    return FQB_OK;
}

// ---- multi-GPU, native (row e) ---------------------------------------------------------------------------------
// Reads shard by batch: global batch b goes to rank b % world, the index is replicated, and the path has no data-path
// collective.  What crosses GPUs:
//  (1) per batch, the 56 bytes bwa_cal_pac_pos_pe carries from batch to batch (position of the drand48 stream,
//      libbwa/bwase.c:33-36 with srand48 once per file, src/BwtMapper.cpp:1817; last_ii, src/BwtMapper.cpp:779-781): the owner
//      of batch b stores them into the mailbox of the owner of b+1 THROUGH PEER MEMORY (NVLink) with a one-thread kernel, and
//      the receiver's pair stage starts with a one-thread kernel that waits for the mailbox's sequence number.  Both are
//      stream-ordered and fit on an SM next to the persistent search kernel, so the hand-off never waits for the GPU to drain
//      (an NCCL send/recv kernel would: it needs a whole CTA's worth of registers while the search grid holds every SM);
//  (2) at the end of the run, the accumulators: one grouped ncclReduce (sum; first-touch contig order: min) onto rank 0, and
//      the variable-size state (marker pile-up entries, distinct PCR-duplicate keys) with exact-size grouped ncclSend/ncclRecv.
}  // extern "C" (reopened below)
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
// NCCL is resolved at first use (dlopen), so the library loads -- and a single GPU works -- where it is not installed
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) { api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL); if (api.lib) break; }
        if (!api.lib) return;
#define FQB_NCCL_SYM(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.lib, "nccl" #f))
        FQB_NCCL_SYM(GetUniqueId); FQB_NCCL_SYM(CommInitRank); FQB_NCCL_SYM(CommInitAll); FQB_NCCL_SYM(CommDestroy); FQB_NCCL_SYM(Reduce);
        FQB_NCCL_SYM(AllGather); FQB_NCCL_SYM(Send); FQB_NCCL_SYM(Recv); FQB_NCCL_SYM(GroupStart); FQB_NCCL_SYM(GroupEnd); FQB_NCCL_SYM(GetErrorString);
#undef FQB_NCCL_SYM
        if (!api.GetUniqueId || !api.CommInitRank || !api.Reduce || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.AllGather) { dlclose(api.lib); api.lib = nullptr; }
    });
    return api.lib ? &api : nullptr;
}
}  // namespace
#define NCCL_CHECK(expr)                                                                                        \
    do {                                                                                                        \
        ncclResult_t r_ = (expr);                                                                               \
        if (r_ != ncclSuccess) { set_error(std::string(#expr) + ": " + (N->GetErrorString ? N->GetErrorString(r_) : "NCCL error")); return FQB_ERR_CUDA; } \
    } while (0)

static_assert(offsetof(fqb_handle::BatchCtl, cur_ii) == 56, "the hand-off state is the first 56 bytes of BatchCtl");
static void comm_release(fqb_handle *h) {
    if (h->nccl) { if (NcclApi *N = nccl_api()) if (N->CommDestroy) N->CommDestroy(h->nccl); h->nccl = nullptr; }
    if (h->ring_next && h->ring_next_ipc) cudaIpcCloseMemHandle(h->ring_next);
    cudaFree(h->ring_inbox);
    h->ring_next = h->ring_inbox = nullptr;
}

extern "C" {
// mailbox of this handle for the hand-off ring; out64 receives its cudaIpcMemHandle_t (for ranks in other processes)
int fqb_comm_ring_handle(fqb_handle *h, uint8_t *out64) {
    if (!h || !out64) { set_error("null argument"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    if (!h->ring_inbox) {
        CU_CHECK(cudaMalloc(&h->ring_inbox, sizeof(fqb_handle::RingBox)));
        CU_CHECK(cudaMemset(h->ring_inbox, 0, sizeof(fqb_handle::RingBox)));
    }
    cudaIpcMemHandle_t ipc;
    static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU_CHECK(cudaIpcGetMemHandle(&ipc, h->ring_inbox));
    memcpy(out64, &ipc, 64);
    return FQB_OK;
}
int fqb_comm_unique_id(uint8_t *out128) {
    NcclApi *N = nccl_api();
    if (!N) { set_error("NCCL (libnccl.so.2) is not available"); return FQB_ERR_IO; }
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    NCCL_CHECK(N->GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return FQB_OK;
}
// one process per GPU: every rank passes the same NCCL id (from rank 0's fqb_comm_unique_id) and all ranks' mailbox handles
// (world x 64 bytes, from fqb_comm_ring_handle), both exchanged by the launcher
int fqb_comm_init(fqb_handle *h, int rank, int world, const uint8_t *nccl_id128, const uint8_t *ring_handles) {
    if (!h || world < 1 || rank < 0 || rank >= world) { set_error("fqb_comm_init: bad rank / world"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    h->comm_rank = rank; h->comm_world = world;
    if (world == 1) return FQB_OK;
    if (!nccl_id128 || !ring_handles) { set_error("fqb_comm_init: the NCCL id and the ring handles are required"); return FQB_ERR_ARG; }
    if (!h->ring_inbox) { set_error("fqb_comm_init: call fqb_comm_ring_handle first"); return FQB_ERR_STATE; }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, ring_handles + (size_t)((rank + 1) % world) * 64, 64);
    void *peer = nullptr;
    CU_CHECK(cudaIpcOpenMemHandle(&peer, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->ring_next = static_cast<fqb_handle::RingBox *>(peer); h->ring_next_ipc = true;
    NcclApi *N = nccl_api();
    if (!N) { set_error("NCCL (libnccl.so.2) is not available"); return FQB_ERR_IO; }
    ncclUniqueId id;
    memcpy(&id, nccl_id128, 128);
    NCCL_CHECK(N->CommInitRank(&h->nccl, world, id, rank));
    return FQB_OK;
}
// the handles of ONE process, rank i = hs[i] (one per GPU; several on one GPU work for the ring, not for NCCL)
int fqb_comm_init_local(fqb_handle **hs, int n) {
    if (!hs || n < 1) { set_error("fqb_comm_init_local: no handles"); return FQB_ERR_ARG; }
    bool distinct = true;
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) if (hs[i]->device == hs[j]->device) distinct = false;
    for (int i = 0; i < n; ++i) {
        uint8_t tmp[64];
        if (int rc = fqb_comm_ring_handle(hs[i], tmp)) return rc;
        hs[i]->comm_rank = i; hs[i]->comm_world = n;
    }
    for (int i = 0; i < n && n > 1; ++i) {
        fqb_handle *nx = hs[(i + 1) % n];
        if (nx->device != hs[i]->device) {
            CU_CHECK(cudaSetDevice(hs[i]->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(nx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return FQB_ERR_CUDA; }
            cudaGetLastError();
        }
        hs[i]->ring_next = nx->ring_inbox; hs[i]->ring_next_ipc = false;
    }
    if (n > 1 && distinct) {
        NcclApi *N = nccl_api();
        if (!N || !N->CommInitAll) { set_error("NCCL (libnccl.so.2) is not available"); return FQB_ERR_IO; }
        std::vector<ncclComm_t> comms(n); std::vector<int> devs(n);
        for (int i = 0; i < n; ++i) devs[i] = hs[i]->device;
        NCCL_CHECK(N->CommInitAll(comms.data(), n, devs.data()));
        for (int i = 0; i < n; ++i) hs[i]->nccl = comms[i];
    }
    return FQB_OK;
}

// fqb_collect_pairs for a sharded run.  global_batch: position of this batch in file order (this rank owns the batches with
// global_batch % world == rank); first_pair: global index of its first pair; is_last: no batch follows in the file.  The pair
// stage starts with the state the owner of global_batch - 1 left (unless this is batch 0 of the file) and hands its own on.
int fqb_collect_pairs_sharded(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2, uint64_t global_batch, uint64_t first_pair, int is_last) {
    if (!h || !h->n_fifo) { set_error("fqb_collect_pairs_sharded: no batch submitted"); return FQB_ERR_STATE; }
    if (h->comm_world > 1 && !h->ring_next) { set_error("fqb_collect_pairs_sharded: call fqb_comm_init first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    const int si = h->fifo[0];
    h->fifo[0] = h->fifo[1]; --h->n_fifo;
    use_set(h, si);
    h->pairs_seen = first_pair;
    cudaStream_t st = h->stream;
    CU_CHECK(cudaStreamWaitEvent(st, h->sets[si].ev_align, 0));
    static const bool tl_on = getenv("FQB_TIMELINE") != nullptr;
    auto tl_mark = [&](cudaStream_t s_) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s_); h->tl_ev.push_back(e); };
    if (tl_on) tl_mark(st);
    const bool ring = h->comm_world > 1;
    const unsigned int recv_seq = ring && global_batch > 0 ? (h->ring_epoch << 20 | (unsigned int)(global_batch & 0xfffffu)) : 0u;
    const unsigned int send_seq = ring && !is_last ? (h->ring_epoch << 20 | (unsigned int)((global_batch + 1) & 0xfffffu)) : 0u;
    int rc = enqueue_pair(h, st, recv_seq, send_seq);
    if (rc) return rc;
    rc = enqueue_sw_refine(h, st);
    if (rc) return rc;
    h->batch_ready = h->align_done = h->pair_done = h->dp_done = true;
    if (rows1 && rows2) { if ((rc = fqb_stage_fetch_rows_async(h, rows1, rows2))) return rc; }
    if (h->stats_open) rc = enqueue_stats(h, st);
    if (!rc) rc = enqueue_status(h, st);
    if (rc) return rc;
    h->stats_done = h->stats_open;
    CU_CHECK(cudaEventRecord(h->sets[si].ev_done, st));
    if (tl_on) tl_mark(st);
    return FQB_OK;
}

// End of a sharded run, called by every rank: the fixed-size accumulators are summed onto rank 0 (one NCCL group), then the
// ranks' pile-up entries and distinct duplicate keys go there with exact-size sends.  Rank 0 ends up with the state of an
// unsharded run (fqb_stats_merge_tables + fqb_stats_finish follow there).  ms_out (optional): device time of the exchange.
int fqb_comm_merge_stats(fqb_handle *h, double *ms_out) {
    if (!h || !h->stats_open) { set_error("fqb_comm_merge_stats: statistics are not open"); return FQB_ERR_STATE; }
    if (ms_out) *ms_out = 0.0;
    if (h->comm_world == 1) return FQB_OK;
    NcclApi *N = nccl_api();
    if (!N || !h->nccl) { set_error("fqb_comm_merge_stats: no NCCL communicator (fqb_comm_init)"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int W = h->comm_world, me = h->comm_rank;
    const size_t nc = h->stabs.contigs.size(), ns = h->stabs.n_sites ? h->stabs.n_sites : 1;
    if (W > kMaxRanks) { set_error("fqb_comm_merge_stats: more than 64 ranks"); return FQB_ERR_LIMIT; }
    {   // per-file counters (FileStatCollector: one per FASTQ pair, src/BwtMapper.cpp:249-254): every rank closes its share of
        // the last file, then the F x 6 table is summed onto rank 0
        if (int rc = close_current_file(h)) return rc;
        const size_t F = h->files.size();
        if (F) {
            std::vector<long long> fc(F * 6);
            for (size_t f = 0; f < F; ++f) {
                const FileCounters &x = h->files[f];
                const long long v[6] = {x.TotalFiltered, x.BwaUnmapped, x.TotalMAPQ, x.TotalRetained, x.NumBase, x.NumRead};
                for (int k = 0; k < 6; ++k) fc[f * 6 + k] = v[k];
            }
            long long *d_fc = nullptr;
            CU_CHECK(cudaMallocAsync(&d_fc, F * 6 * 8, st));
            CU_CHECK(cudaMemcpyAsync(d_fc, fc.data(), F * 6 * 8, cudaMemcpyHostToDevice, st));
            NCCL_CHECK(N->Reduce(d_fc, d_fc, F * 6, ncclInt64, ncclSum, 0, h->nccl, st));
            CU_CHECK(cudaMemcpyAsync(fc.data(), d_fc, F * 6 * 8, cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaStreamSynchronize(st));
            CU_CHECK(cudaFreeAsync(d_fc, st));
            if (me == 0)
                for (size_t f = 0; f < F; ++f) {
                    FileCounters &x = h->files[f];
                    x.TotalFiltered = fc[f * 6]; x.BwaUnmapped = fc[f * 6 + 1]; x.TotalMAPQ = fc[f * 6 + 2]; x.TotalRetained = fc[f * 6 + 3];
                    x.NumBase = fc[f * 6 + 4]; x.NumRead = fc[f * 6 + 5];
                }
        }
        h->files_merged = true;
    }
    // what every rank holds: its pile-up entries and, per owner rank, its distinct duplicate keys (this first collective also
    // absorbs the skew between the ranks, so that the events below time the exchange itself)
    uint64_t n_tup_mine = 0;
    { int rc = fqb_stats_var_count(h, 0, &n_tup_mine); if (rc) return rc; }
    const int RW = W + 1;                              // row: [pile-up entries, keys for owner 0 .. W-1]
    uint64_t *d_cnt = nullptr;
    CU_CHECK(cudaMallocAsync(&d_cnt, (size_t)(W + 2) * RW * 8, st));
    uint64_t *d_row = d_cnt + (size_t)W * RW, *d_cur = d_row + RW;
    CU_CHECK(cudaMemsetAsync(d_row, 0, (size_t)2 * RW * 8, st));
    CU_CHECK(cudaMemcpyAsync(d_row, &n_tup_mine, 8, cudaMemcpyHostToDevice, st));
    dup_part_count_kernel<<<1024, 256, 0, st>>>(h->d_dup_keys, h->dup_cap, W, reinterpret_cast<unsigned long long *>(d_row + 1));
    NCCL_CHECK(N->AllGather(d_row, d_cnt, RW, ncclUint64, h->nccl, st));
    std::vector<uint64_t> cnt((size_t)W * RW);
    CU_CHECK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)W * RW * 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    auto keys_of = [&](int from, int to) { return cnt[(size_t)from * RW + 1 + to]; };
    uint64_t n_mine = 0, n_recv = 0;
    for (int d = 0; d < W; ++d) { n_mine += keys_of(me, d); n_recv += keys_of(d, me); }
    cudaEvent_t e0, e1;
    CU_CHECK(cudaEventCreate(&e0)); CU_CHECK(cudaEventCreate(&e1));
    CU_CHECK(cudaEventRecord(e0, st));
    // ---- duplicate keys: all-to-all by owner, every owner counts what it sees more than once
    {
        unsigned long long tab = 1ull << 20;
        while (tab < 2 * n_recv) tab <<= 1;
        const size_t need = (size_t)n_mine + n_recv + tab;
        if (need > h->cap_keys_imp) {
            // kept for the life of the handle and sized generously (512 MB at least): allocating inside the exchange costs 10+ ms
            cudaFree(h->d_keys_imp);
            h->cap_keys_imp = need + need / 2 < (64u << 20) ? (64u << 20) : need + need / 2;
            CU_CHECK(cudaMalloc(&h->d_keys_imp, h->cap_keys_imp * 8));
        }
        unsigned long long *send = h->d_keys_imp, *recv = send + n_mine, *table = recv + n_recv;
        dup_part_scatter_kernel<<<1024, 256, 0, st>>>(h->d_dup_keys, h->dup_cap, W, reinterpret_cast<unsigned long long *>(d_row + 1),
                                                      reinterpret_cast<unsigned long long *>(d_cur), send);
        CU_CHECK(cudaMemsetAsync(table, 0, (size_t)tab * 8, st));
        NCCL_CHECK(N->GroupStart());
        uint64_t so = 0, ro = 0;
        for (int d = 0; d < W; ++d) {
            const uint64_t ns_ = keys_of(me, d), nr = keys_of(d, me);
            if (d == me) { if (ns_) CU_CHECK(cudaMemcpyAsync(recv + ro, send + so, ns_ * 8, cudaMemcpyDeviceToDevice, st)); }
            else {
                if (ns_) NCCL_CHECK(N->Send(send + so, ns_ * 8, ncclUint8, d, h->nccl, st));
                if (nr) NCCL_CHECK(N->Recv(recv + ro, nr * 8, ncclUint8, d, h->nccl, st));
            }
            so += ns_; ro += nr;
        }
        NCCL_CHECK(N->GroupEnd());
        if (n_recv) dup_cross_kernel<<<(unsigned)((n_recv + 255) / 256), 256, 0, st>>>(table, tab - 1, recv, n_recv, h->d_emp + (4 * 256 + 4096));
        h->n_launches += 3;
    }
    CU_CHECK(cudaFreeAsync(d_cnt, st));
    cudaEvent_t e_keys;
    CU_CHECK(cudaEventCreate(&e_keys));
    CU_CHECK(cudaEventRecord(e_keys, st));
    // ---- fixed-size accumulators (NumPCRDup now includes this rank's share of the cross-rank duplicates)
    NCCL_CHECK(N->GroupStart());
    NCCL_CHECK(N->Reduce(h->d_depth, h->d_depth, ns * 3, ncclUint32, ncclSum, 0, h->nccl, st));
    NCCL_CHECK(N->Reduce(h->d_emp, h->d_emp, (size_t)kEmpWords, ncclUint64, ncclSum, 0, h->nccl, st));
    NCCL_CHECK(N->Reduce(h->d_contig_ctr, h->d_contig_ctr, nc * 4, ncclUint32, ncclSum, 0, h->nccl, st));
    NCCL_CHECK(N->Reduce(h->d_contig_ctr + nc * 4, h->d_contig_ctr + nc * 4, nc, ncclUint32, ncclMin, 0, h->nccl, st));
    NCCL_CHECK(N->GroupEnd());
    cudaEvent_t e_red;
    CU_CHECK(cudaEventCreate(&e_red));
    CU_CHECK(cudaEventRecord(e_red, st));
    // ---- pile-up entries to rank 0, exact sizes
    if (me != 0) {
        const bool direct = h->tuples_host.empty() && h->n_tuples_imp == 0;      // usually on the device in one piece
        void *tup = direct ? (void *)h->d_tuples : nullptr;
        if (!direct && n_tup_mine) {
            CU_CHECK(cudaMallocAsync(&tup, n_tup_mine * sizeof(PileupTuple), st));
            int rc = fqb_stats_var_export(h, 0, tup, n_tup_mine);
            if (rc) return rc;
        }
        if (n_tup_mine) NCCL_CHECK(N->Send(tup, n_tup_mine * sizeof(PileupTuple), ncclUint8, 0, h->nccl, st));
        if (!direct && tup) CU_CHECK(cudaFreeAsync(tup, st));
    } else {
        uint64_t n_tup = 0;
        for (int r = 1; r < W; ++r) n_tup += cnt[(size_t)r * RW];
        if (h->n_tuples_imp + n_tup > h->cap_tuples_imp) {
            // grown with headroom and kept: allocating inside the exchange costs tens of milliseconds
            size_t cap = (h->n_tuples_imp + n_tup) * 2;
            if (cap < (4u << 20)) cap = 4u << 20;
            PileupTuple *nb = nullptr;
            CU_CHECK(cudaMalloc(&nb, cap * sizeof(PileupTuple)));
            if (h->n_tuples_imp) CU_CHECK(cudaMemcpyAsync(nb, h->d_tuples_imp, h->n_tuples_imp * sizeof(PileupTuple), cudaMemcpyDeviceToDevice, st));
            CU_CHECK(cudaStreamSynchronize(st));
            cudaFree(h->d_tuples_imp);
            h->d_tuples_imp = nb; h->cap_tuples_imp = cap;
        }
        NCCL_CHECK(N->GroupStart());
        uint64_t to = h->n_tuples_imp;
        for (int r = 1; r < W; ++r) {
            const uint64_t c = cnt[(size_t)r * RW];
            if (c) NCCL_CHECK(N->Recv(h->d_tuples_imp + to, c * sizeof(PileupTuple), ncclUint8, r, h->nccl, st));
            to += c;
        }
        NCCL_CHECK(N->GroupEnd());
        h->n_tuples_imp += n_tup;
    }
    CU_CHECK(cudaEventRecord(e1, st));
    CU_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (getenv("FQB_COMM_DEBUG")) {
        float m1 = 0.f, m2 = 0.f;
        cudaEventElapsedTime(&m1, e0, e_keys); cudaEventElapsedTime(&m2, e_keys, e_red);
        fprintf(stderr, "rank %d merge: duplicate keys %.2f ms (%llu sent, %llu owned), reduce %.2f ms, pile-up entries %.2f ms (%llu)\n", me, m1,
                (unsigned long long)n_mine, (unsigned long long)n_recv, m2, ms - m1 - m2, (unsigned long long)n_tup_mine);
    }
    cudaEventDestroy(e_red); cudaEventDestroy(e_keys);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return FQB_OK;
}

// result rows of the last completed stage: rows[e][i] = end e of pair i
}  // extern "C" (reopened below)
// out[end][pair] = in[pair][end], moved as 16-byte words (sizeof(fqb_read_t) = 6 x 16)
static __global__ void split_rows_kernel(const uint4 *in, uint4 *out, size_t n_pairs) {
    constexpr size_t kW = sizeof(fqb_read_t) / 16;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs * 2 * kW) return;
    const size_t row = t / kW, w = t % kW, pair = row >> 1, end = row & 1;
    out[(end * n_pairs + pair) * kW + w] = in[t];
}
extern "C" {
int fqb_stage_fetch_rows(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out) {
    if (!h || !h->pair_done) { set_error("fqb_stage_fetch_rows: run fqb_stage_pair first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    const size_t np = (size_t)h->n_reads / 2;
    CU_CHECK(cudaStreamWaitEvent(h->stream, h->ev_rows[2], 0));          // a pending asynchronous fetch still reads d_rows_split
    if (np) {
        // rows are interleaved by end on the device (r = 2*pair + end); split them there so that the two
        // device-to-host copies are contiguous (strided 96-byte copies crawl over PCIe)
        split_rows_kernel<<<(unsigned)((np * 2 * (sizeof(fqb_read_t) / 16) + 255) / 256), 256, 0, h->stream>>>(
            reinterpret_cast<const uint4 *>(h->d_rows), reinterpret_cast<uint4 *>(h->d_rows_split), np);
        ++h->n_launches;
        CU_CHECK(cudaMemcpyAsync(rows1, h->d_rows_split, np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, h->stream));
        CU_CHECK(cudaMemcpyAsync(rows2, h->d_rows_split + np, np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (int rc = sync_and_check(h)) return rc;
    if (ii_out) *ii_out = h->cur_ii;
    return FQB_OK;
}

// The same copy, asynchronous: the rows of the resident batch are split and copied on a second stream while the caller
// goes on to the next batch (the engine's own stream only runs the 20-us split).
// fqb_rows_wait blocks until the destination buffers of the last fqb_stage_fetch_rows_async are complete.
int fqb_stage_fetch_rows_async(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2) {
    if (!h || !h->pair_done || !rows1 || !rows2) { set_error("fqb_stage_fetch_rows_async: run fqb_stage_pair first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    const size_t np = (size_t)h->n_reads / 2;
    // the split runs on the main stream itself (high priority: it finds SM room next to the resident search grid at once; on
    // the copy-out stream it queued behind that grid and the next batch's later stages waited for it), the two contiguous
    // copies on the copy-out stream.  The main stream only waits for the PREVIOUS batch's copies, which still read d_rows_split
    CU_CHECK(cudaStreamWaitEvent(h->stream, h->ev_rows[2], 0));
    if (np) {
        split_rows_kernel<<<(unsigned)((np * 2 * (sizeof(fqb_read_t) / 16) + 255) / 256), 256, 0, h->stream>>>(
            reinterpret_cast<const uint4 *>(h->d_rows), reinterpret_cast<uint4 *>(h->d_rows_split), np);
        ++h->n_launches;
    }
    CU_CHECK(cudaEventRecord(h->ev_rows[0], h->stream));
    CU_CHECK(cudaStreamWaitEvent(h->d2h_stream, h->ev_rows[0], 0));
    if (np) {
        CU_CHECK(cudaMemcpyAsync(rows1, h->d_rows_split, np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, h->d2h_stream));
        CU_CHECK(cudaMemcpyAsync(rows2, h->d_rows_split + np, np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, h->d2h_stream));
    }
    CU_CHECK(cudaEventRecord(h->ev_rows[2], h->d2h_stream));
    return FQB_OK;
}
int fqb_rows_wait(fqb_handle *h) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaEventSynchronize(h->ev_rows[2]));
    if (h->tl_base && !h->tl_ev.empty()) {
        cudaDeviceSynchronize();
        fprintf(stderr, "timeline (ms since the first submit; events in enqueue order: per submit [align start, align end], per collect [later start, later end])\n");
        for (size_t k = 0; k < h->tl_ev.size(); ++k) { float ms = 0; cudaEventElapsedTime(&ms, h->tl_base, h->tl_ev[k]); fprintf(stderr, " %.2f", ms); cudaEventDestroy(h->tl_ev[k]); }
        fprintf(stderr, "\n");
        h->tl_ev.clear(); cudaEventDestroy(h->tl_base); h->tl_base = nullptr;
    }
    return sync_and_check(h);            // also reports what the device flagged for the batches completed since the last check
}

// restart the per-file state: srand48(bns->seed) and last_ii (PairEndMapper runs once per FASTQ pair)
int fqb_reset_stream(fqb_handle *h) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    h->rng_calls = 0;
    h->last_ii.avg = h->last_ii.std = -1.0; h->last_ii.low = h->last_ii.high = h->last_ii.high_bayesian = 0;
    ++h->ring_epoch;
    return push_stream_state(h);
}

int fqb_stage_fetch_prep(fqb_handle *h, int32_t *len, int32_t *full_len, uint8_t *filtered, uint8_t *codes, int32_t codes_stride) {
    if (!h || !h->batch_ready) { set_error("no batch loaded"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    size_t n = (size_t)h->n_reads;
    if (len) CU_CHECK(cudaMemcpy(len, h->bv.len, n * 4, cudaMemcpyDeviceToHost));
    if (full_len) CU_CHECK(cudaMemcpy(full_len, h->bv.full_len, n * 4, cudaMemcpyDeviceToHost));
    if (filtered) CU_CHECK(cudaMemcpy(filtered, h->bv.filtered, n, cudaMemcpyDeviceToHost));
    if (codes) CU_CHECK(cudaMemcpy2D(codes, (size_t)codes_stride, h->bv.codes, (size_t)h->lpad, (size_t)h->stride, n, cudaMemcpyDeviceToHost));
    return FQB_OK;
}

// hits per read r = 2*pair+end, padded to `cap` rows; n_aln[r] is the true count
int fqb_stage_fetch_aln(fqb_handle *h, int32_t cap, fqb_aln_t *out, int32_t *n_aln) {
    if (!h || !h->batch_ready || cap < 1) { set_error("no batch loaded"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    size_t n = (size_t)h->n_reads;
    std::vector<Hit> fast(n * kAlnCapFast);
    std::vector<int32_t> cnt(n), slot(n);
    CU_CHECK(cudaMemcpy(fast.data(), h->d_aln, fast.size() * sizeof(Hit), cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemcpy(cnt.data(), h->d_naln, n * 4, cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemcpy(slot.data(), h->d_spill_slot, n * 4, cudaMemcpyDeviceToHost));
    std::vector<Hit> row(kAlnCapSlow);
    memset(out, 0, n * (size_t)cap * sizeof(fqb_aln_t));
    for (size_t r = 0; r < n; ++r) {
        const Hit *src = &fast[r * kAlnCapFast];
        int have = kAlnCapFast;
        if (slot[r] >= 0) {
            CU_CHECK(cudaMemcpy(row.data(), h->d_aln_big + (size_t)slot[r] * kAlnCapSlow, sizeof(Hit) * kAlnCapSlow, cudaMemcpyDeviceToHost));
            src = row.data(); have = kAlnCapSlow;
        }
        int m = cnt[r] < cap ? cnt[r] : cap;
        if (m > have) m = have;
        for (int j = 0; j < m; ++j) {
            fqb_aln_t &o = out[r * (size_t)cap + j];
            o.k = src[j].k; o.l = src[j].l; o.score = src[j].score;
            o.n_mm = src[j].n_mm; o.n_gapo = src[j].n_gapo; o.n_gape = src[j].n_gape; o.a = src[j].a;
        }
        if (n_aln) n_aln[r] = cnt[r];
    }
    return FQB_OK;
}

// since creation: [0] stack pops, [1] rank-query pairs issued by the search kernel, [2] reference-equivalent
// occ-block touches N_blk of bwt_cal_width + bwt_match_gap; [3] overflow reads of the last batch
int fqb_stage_counters(fqb_handle *h, uint64_t *out4) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    CU_CHECK(cudaStreamSynchronize(h->stream));
    unsigned long long c[3];
    uint32_t ov = 0;
    CU_CHECK(cudaMemcpy(c, h->d_counters, 24, cudaMemcpyDeviceToHost));
    if (getenv("FQB_KSTATS_PRINT")) {
        unsigned long long k[16];
        cudaMemcpy(k, h->d_counters, 16 * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "kstats iter %llu step %llu exh %llu hist", k[4], k[5], k[6]);
        for (int q = 0; q < 9; ++q) fprintf(stderr, " %llu", k[7 + q]);
        fprintf(stderr, "\n");
        std::vector<unsigned long long> tl(512);
        cudaMemcpy(tl.data(), h->d_counters, 512 * 8, cudaMemcpyDeviceToHost);
        if (tl[16]) {
            fprintf(stderr, "timeline dry %.2f ms end %.2f ms\n  bin(0.25ms): warps_left trips live/trip\n", (tl[17] > tl[16] ? (tl[17] - tl[16]) * 1e-6 : 0.0), (tl[18] - tl[16]) * 1e-6);
            for (int q = 0; q < 128; ++q) if (tl[32 + q] || tl[160 + q]) fprintf(stderr, "  %3d %6llu %8llu %5.1f\n", q, tl[32 + q], tl[160 + q], tl[160 + q] ? (double)tl[288 + q] / tl[160 + q] : 0.0);
            cudaMemset(h->d_counters + 16, 0, (512 - 16) * 8);
        }
    }
    CU_CHECK(cudaMemcpy(&ov, h->d_ctrs + 2, 4, cudaMemcpyDeviceToHost));
    out4[0] = c[0]; out4[1] = c[1]; out4[2] = c[2]; out4[3] = ov;
    return FQB_OK;
}

int fqb_stage_fetch_rows(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out);
int fqb_stage_pair(fqb_handle *h);
int fqb_stage_sw_refine(fqb_handle *h);

// ---- BAM emission (row f1): SetSamFileHeader / SetSamRecord / BamIO.writeRecord (src/BwtMapper.cpp:947-1264, 2068-2074) ----
int fqb_bam_open(fqb_handle *h, const char *path, const char *rg_line) {
    if (!h || !path) { set_error("null argument"); return FQB_ERR_ARG; }
    if (!h->stats_open) { set_error("fqb_bam_open: call fqb_stats_open first (it loads the reference's contig list)"); return FQB_ERR_STATE; }
    if (h->bam_open) { set_error("a BAM file is already open on this handle"); return FQB_ERR_STATE; }
    std::string err, header;
    if (!bam_prepare(h->hidx, h->gopt, h->stabs.genome_contigs, rg_line ? rg_line : "", h->bam_ctx, header, err)) { set_error(err); return FQB_ERR_IO; }
    if (!h->bam.open(path, err)) { set_error(err); return FQB_ERR_IO; }
    h->bam.write(header.data(), header.size());
    h->bam_open = true;
    return FQB_OK;
}

int fqb_bam_emit2(fqb_handle *h, const char *names, const char *names2, int32_t name_stride, const uint8_t *bases1, const uint8_t *quals1,
                  const uint8_t *bases2, const uint8_t *quals2, int32_t stride) {
    if (!names2) names2 = names;
    fqb_handle *const o_ = (h && h->bam_owner) ? h->bam_owner : h;     // whose file, record context and read-buffer history
    if (!h || !o_->bam_open) { set_error("fqb_bam_emit: no BAM file open"); return FQB_ERR_STATE; }
    if (!h->dp_done || (h->stats_open && !h->stats_done)) { set_error("fqb_bam_emit: the batch must be through fqb_stage_sw_refine and fqb_stage_stats"); return FQB_ERR_STATE; }
    if (!bases1 || !quals1 || stride < 1 || (!h->single_end && (!bases2 || !quals2))) { set_error("fqb_bam_emit: the reads of the batch are required"); return FQB_ERR_ARG; }
    CU_CHECK(cudaSetDevice(h->device));
    if (int rc = drain_post(h, false, true)) return rc;      // the previous batch's records are out; fqb_stats_emit's host phase of THIS batch may still run
    cudaStream_t st = h->stream;
    const size_t np = (size_t)h->n_reads / 2;
    const auto t_begin = std::chrono::steady_clock::now();
    // the other hits of reads that keep a multi list (XA): positions and CIGARs come from the device
    if (h->multi_cap < (uint32_t)h->cap_reads) {
        cudaFree(h->d_multi_out); cudaFree(h->d_multi_list); cudaFree(h->d_multi_ctr);
        h->multi_cap = (uint32_t)h->cap_reads;                                    // multi lists are rare (n_occ <= n_multi + 1); one slot per read is ample
        CU_CHECK(cudaMalloc(&h->d_multi_out, (size_t)h->multi_cap * sizeof(MultiOut)));
        CU_CHECK(cudaMalloc(&h->d_multi_list, (size_t)h->cap_reads * 2 * 4));
        CU_CHECK(cudaMalloc(&h->d_multi_ctr, 4 * 4));
    }
    DpView dv;
    dv.n_reads = h->n_reads; dv.lpad = h->lpad; dv.codes = h->bv.codes; dv.pac = h->d_pac; dv.l_pac = h->hidx.l_pac; dv.rows = h->d_rows;
    MultiView mv;
    mv.aln = h->d_aln; mv.aln_cap = kAlnCapFast; mv.aln_big = h->d_aln_big; mv.aln_big_cap = kAlnCapSlow; mv.spill_slot = h->d_spill_slot; mv.n_aln = h->d_naln;
    mv.bwt[0] = h->dbwt[0]; mv.bwt[1] = h->dbwt[1];
    CU_CHECK(cudaMemsetAsync((h->d_dpctr + 8), 0, 4, st));
    launch_multi(dv, mv, h->dp_pool, h->d_multi_list, h->d_multi_ctr, h->d_multi_out, h->multi_cap, (h->d_dpctr + 8), st);
    h->n_launches += 2;
    uint32_t ctr[2] = {0, 0}, derr = 0;
    CU_CHECK(cudaMemcpyAsync(ctr, h->d_multi_ctr, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(&derr, (h->d_dpctr + 8), 4, cudaMemcpyDeviceToHost, st));
    if (np > h->h_bam_rows_cap) {
        cudaFreeHost(h->h_bam_rows); h->h_bam_rows = nullptr; h->h_bam_rows_cap = 0;
        CU_CHECK(cudaMallocHost(&h->h_bam_rows, 2 * np * sizeof(fqb_read_t)));
        h->h_bam_rows_cap = np;
    }
    CU_CHECK(cudaMemcpyAsync(h->h_bam_rows, h->d_rows, 2 * np * sizeof(fqb_read_t), cudaMemcpyDeviceToHost, st));
    const bool have_ps = h->stats_done && h->d_pstat;
    const uint64_t first = h->pairs_seen - (h->stats_done ? np : 0);
    if (have_ps && h->h_pstat_first != first) {              // not staged by fqb_stats_emit already
        if (int rc = drain_post(h, true, true)) return rc;
        if (np > h->h_rows_cap) {
            cudaFreeHost(h->h_rows); cudaFreeHost(h->h_pstat);
            h->h_rows = nullptr; h->h_pstat = nullptr; h->h_rows_cap = 0;
            CU_CHECK(cudaMallocHost(&h->h_rows, 2 * np * sizeof(fqb_read_t)));
            CU_CHECK(cudaMallocHost(&h->h_pstat, np * sizeof(PairStat)));
            h->h_rows_cap = np;
        }
        CU_CHECK(cudaMemcpyAsync(h->h_pstat, h->d_pstat, np * sizeof(PairStat), cudaMemcpyDeviceToHost, st));
        h->h_pstat_first = first;
    }
    CU_CHECK(cudaStreamSynchronize(st));
    if (derr) { set_error("multi-hit list: capacity exceeded for read " + std::to_string(derr - 1)); return FQB_ERR_LIMIT; }
    auto mo_p = std::make_shared<std::vector<MultiOut>>(ctr[1]);
    std::vector<MultiOut> &mo = *mo_p;
    if (ctr[1]) CU_CHECK(cudaMemcpy(mo.data(), h->d_multi_out, (size_t)ctr[1] * sizeof(MultiOut), cudaMemcpyDeviceToHost));
    std::sort(mo.begin(), mo.end(), [](const MultiOut &a, const MultiOut &b) { return a.read != b.read ? a.read < b.read : a.j < b.j; });
    auto xa_p = std::make_shared<std::vector<XaHit>>(mo.size());
    for (size_t i = 0; i < mo.size(); ++i) {
        XaHit &x = (*xa_p)[i];
        x.pos = mo[i].pos; x.strand = mo[i].strand; x.gap = mo[i].gap; x.mm = mo[i].mm; x.has_cigar = mo[i].has_cigar; x.n_cigar = mo[i].n_cigar;
        memcpy(x.cigar, mo[i].cigar, sizeof x.cigar);
    }
    const auto t_dev = std::chrono::steady_clock::now();
    const int par = (int)(o_->bam_batches & 1);
    if (!h->single_end) {
        if (o_->rseq_stride != (size_t)stride) { for (auto &a : o_->rseq_shadow) for (auto &b : a) b.clear(); o_->rseq_stride = (size_t)stride; }
        for (int e = 0; e < 2; ++e) if (o_->rseq_shadow[par][e].size() < (size_t)(FQB_BATCH_PAIRS > np ? FQB_BATCH_PAIRS : np) * stride)
            o_->rseq_shadow[par][e].resize((size_t)(FQB_BATCH_PAIRS > np ? FQB_BATCH_PAIRS : np) * stride, 0);
    }
    ++o_->bam_batches;
    const unsigned n_multi_hits = ctr[1];
    // host phase: records are formatted by several host threads over contiguous slices of the batch and handed to the BGZF
    // writer in order.  names / bases / quals must stay valid until the next fqb_stats_emit / fqb_bam_emit / fqb_stats_finish /
    // fqb_bam_close call on this handle returns.
    auto host_phase = [=]() {
    const std::vector<MultiOut> &mo = *mo_p;
    const std::vector<XaHit> &xa = *xa_p;
    auto xa_of = [&](uint32_t r, int &n) -> const XaHit * {
        auto lo = std::lower_bound(mo.begin(), mo.end(), r, [](const MultiOut &a, uint32_t v) { return a.read < v; });
        auto hi = lo;
        while (hi != mo.end() && hi->read == r) ++hi;
        n = (int)(hi - lo);
        return n ? &xa[(size_t)(lo - mo.begin())] : nullptr;
    };
    unsigned nthr = std::thread::hardware_concurrency();
    if (nthr < 1) nthr = 1;
    if (nthr > 16) nthr = 16;
    if (np < 4096) nthr = 1;
    std::vector<std::string> parts(nthr);
    auto work = [&](unsigned t) {
        std::string &o = parts[t];
        const size_t lo = np * t / nthr, hi = np * (t + 1) / nthr;
        o.reserve((hi - lo) * 2 * (size_t)(160 + 3 * stride / 2));
        char buf[64];
        std::string nm, nm2;
        const uint8_t *nt4 = nt4_table();
        for (size_t i = lo; i < hi; ++i) {
            const fqb_read_t &p = h->h_bam_rows[2 * i], &q = h->h_bam_rows[2 * i + 1];
            const uint8_t *rs[2] = {nullptr, nullptr};
            if (!h->single_end) {                      // what bwa_read_seq_with_hash_dev / expand_seq wrote into the slot's rseq
                const fqb_read_t *rw[2] = {&p, &q};
                const uint8_t *bs[2] = {bases1 + i * (size_t)stride, bases2 + i * (size_t)stride};
                for (int e = 0; e < 2; ++e) {
                    uint8_t *dst = o_->rseq_shadow[par][e].data() + i * (size_t)stride;
                    if (!rw[e]->filtered)
                        for (int k = 0; k < rw[e]->clip_len; ++k) { const uint8_t c = nt4[bs[e][rw[e]->clip_len - 1 - k]]; dst[k] = c < 4 ? (uint8_t)(3 - c) : (uint8_t)4; }
                    rs[e] = dst;
                }
            }
            // skip decisions use the types the reads had before AddAlignment's bridge check (kept in the pair record)
            if (have_ps ? (h->h_pstat[i].both_filtered || h->h_pstat[i].both_unmapped)
                        : ((p.filtered && q.filtered) || (p.type == kTypeNoMatch && q.type == kTypeNoMatch))) continue;
            const char *name;
            if (names) { nm.assign(names + i * (size_t)name_stride, strnlen(names + i * (size_t)name_stride, (size_t)name_stride)); name = nm.c_str(); }
            else { snprintf(buf, sizeof buf, "r%011llu", (unsigned long long)(first + i)); name = buf; }
            const char *name_q = name;             // the mates' names may differ (each record carries its own read's, SetSamRecord)
            if (names2 && names2 != names) { nm2.assign(names2 + i * (size_t)name_stride, strnlen(names2 + i * (size_t)name_stride, (size_t)name_stride)); name_q = nm2.c_str(); }
            int n0 = 0, n1 = 0;
            const XaHit *x0 = p.n_multi ? xa_of((uint32_t)(2 * i), n0) : nullptr, *x1 = q.n_multi ? xa_of((uint32_t)(2 * i + 1), n1) : nullptr;
            if (h->single_end) bam_append_single(o_->bam_ctx, p, name, bases1 + i * (size_t)stride, quals1 + i * (size_t)stride, x0, n0, o);
            else bam_append_pair(o_->bam_ctx, p, q, name, name_q, bases1 + i * (size_t)stride, quals1 + i * (size_t)stride, bases2 + i * (size_t)stride,
                                 quals2 + i * (size_t)stride, x0, n0, x1, n1, rs[0], rs[1], o);
        }
    };
    if (nthr == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    const auto t_fmt = std::chrono::steady_clock::now();
    size_t total = 0;
    for (auto &o : parts) { total += o.size(); o_->bam.write_owned(std::move(o)); }
    if (getenv("FQB_BAM_DEBUG")) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "bam_emit: %zu pairs, %u multi hits, %zu bytes, threads %u; device+copies %.1f ms, format %.1f ms, hand-over %.1f ms\n", np, n_multi_hits, total, nthr,
                std::chrono::duration<double, std::milli>(t_dev - t_begin).count(), std::chrono::duration<double, std::milli>(t_fmt - t_dev).count(),
                std::chrono::duration<double, std::milli>(t_end - t_fmt).count());
    }
    };
    if (emit_inline() || h->bam_owner) { host_phase(); return drain_post(h, true, true); }      // records of several handles must reach a shared file in call order
    h->post_bam = std::async(std::launch::async, host_phase);
    return FQB_OK;
}

int fqb_bam_emit(fqb_handle *h, const char *names, int32_t name_stride, const uint8_t *bases1, const uint8_t *quals1,
                 const uint8_t *bases2, const uint8_t *quals2, int32_t stride) {
    return fqb_bam_emit2(h, names, nullptr, name_stride, bases1, quals1, bases2, quals2, stride);
}

// sharded run inside one process: the records this handle formats go, in call order, to `owner`'s BAM file
int fqb_bam_attach(fqb_handle *h, fqb_handle *owner) {
    if (!h || !owner || !owner->bam_open) { set_error("fqb_bam_attach: the owner has no BAM file open"); return FQB_ERR_STATE; }
    h->bam_owner = owner == h ? nullptr : owner;
    return FQB_OK;
}

int fqb_bam_close(fqb_handle *h) {
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    if (!h->bam_open) return FQB_OK;
    const int rc = drain_post(h, true, true);
    h->bam_open = false;
    std::string err;
    if (!h->bam.close(err)) { set_error(err); return FQB_ERR_IO; }
    return rc;
}

// The whole per-batch body of BwtMapper::PairEndMapper up to (not including) the statistics loop.
// Upload the NEXT batch on the copy stream while the current one is being processed; the following
// fqb_align_pairs / fqb_stage_load call with the same host pointers and shape picks it up without copying.
int fqb_prefetch_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                       const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2) {
    if (int rc = check_shape(h, n_pairs, stride, bases1, quals1, bases2, quals2)) return rc;
    CU_CHECK(cudaSetDevice(h->device));
    if (2 * n_pairs > h->cap_reads || stride > h->stride_cap || h->n_fifo) return FQB_OK;      // buffers not sized yet: the load will copy
    // into the staging arrays of the set that is not current, unless they still hold a prefetched batch nobody has
    // loaded yet (the documented call order is prefetch(n+1) before align(n)): then the current set's, free once its
    // prep_kernel has run
    int si = (h->cur + 1) % kSets;
    if (h->sets[si].pre_valid) si = h->cur;
    fqb_handle::BatchSet &B = h->sets[si];
    if (B.pre_valid) return FQB_OK;                  // both hold unconsumed uploads: nothing to do, the load will copy
    const uint8_t *src[4] = {bases1, quals1, bases2, quals2};
    const int32_t *lsrc[2] = {lens1, lens2};
    const size_t bytes = (size_t)n_pairs * stride;
    CU_CHECK(cudaStreamWaitEvent(h->copy_stream, B.ev_free, 0));
    for (int i = 0; i < 4; ++i) if (src[i]) CU_CHECK(cudaMemcpyAsync(B.d_in[i], src[i], bytes, cudaMemcpyHostToDevice, h->copy_stream));
    for (int i = 0; i < 2; ++i)
        if (lsrc[i]) CU_CHECK(cudaMemcpyAsync(B.d_lens_in[i], lsrc[i], (size_t)n_pairs * 4, cudaMemcpyHostToDevice, h->copy_stream));
    CU_CHECK(cudaEventRecord(B.ev_in, h->copy_stream));
    B.pre_valid = true; B.pre_pairs = n_pairs; B.pre_stride = stride;
    for (int i = 0; i < 4; ++i) B.pre_key[i] = src[i];
    return FQB_OK;
}
uint64_t fqb_prefetch_hits(const fqb_handle *h) { return h ? h->prefetch_hits : 0; }

static int align_pairs_impl(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                            const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2,
                            fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out, int32_t packed_stride) {
    int rc = stage_load_impl(h, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, 0, packed_stride);
    if (rc) return rc;
    // the whole chain is enqueued at once; the only wait is for the result
    if ((rc = enqueue_align(h, h->cur, h->stream))) return rc;
    if ((rc = enqueue_pair(h, h->stream))) return rc;
    if ((rc = enqueue_sw_refine(h, h->stream))) return rc;
    if ((rc = enqueue_status(h, h->stream))) return rc;
    h->align_done = h->pair_done = h->dp_done = true; h->stats_done = false;
    if (rows1 && rows2) return fqb_stage_fetch_rows(h, rows1, rows2, ii_out);
    if ((rc = sync_and_check(h))) return rc;
    if (ii_out) *ii_out = h->cur_ii;
    return FQB_OK;
}
int fqb_align_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                    const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2,
                    fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out) {
    return align_pairs_impl(h, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, rows1, rows2, ii_out, 0);
}
// the same batch in the packed input form (fqb_pack_reads / fqb_feeder_fill_packed): 2-bit bases, not-ACGT flag in bit 7 of the qualities
int fqb_align_pairs_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride, const uint8_t *packed1, const uint8_t *quals1,
                           const int32_t *lens1, const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2,
                           fqb_read_t *rows1, fqb_read_t *rows2, fqb_isize_t *ii_out) {
    if (packed_stride <= 0) { set_error("fqb_align_pairs_packed: packed_stride must be positive"); return FQB_ERR_ARG; }
    return align_pairs_impl(h, n_pairs, stride, packed1, quals1, lens1, packed2, quals2, lens2, rows1, rows2, ii_out, packed_stride);
}

// ---- pipelined form: submit batch n+1, then collect batch n ---------------------------------------------------
// fqb_submit_pairs uploads a batch and enqueues its align stage (a1-a5) on the align stream, into the batch set that is
// free; it returns at once.  fqb_collect_pairs takes the OLDEST submitted batch through pairing, mate rescue, refinement
// and -- when statistics are open -- StatCollector's accumulation on the main stream, and starts the copy of its result
// rows; it does not wait either.  The align stage of batch n+1 therefore runs on the GPU next to the later stages of
// batch n, and nothing in the loop blocks the host: fqb_rows_wait (or any call that returns data) is the only wait.
static int submit_impl(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                       const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device, int32_t packed_stride) {
    if (int rc = check_shape(h, n_pairs, stride, bases1, quals1, bases2, quals2, packed_stride)) return rc;
    if (h->n_fifo >= 2) { set_error("fqb_submit_pairs: two batches are already in flight; collect one first"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    if (2 * n_pairs > h->cap_reads || stride > h->stride_cap) {
        if (h->n_fifo) { set_error("fqb_submit_pairs: the batch buffers must grow while a batch is in flight; collect it first"); return FQB_ERR_STATE; }
        int rc = ensure_batch(h, 2 * n_pairs, stride);
        if (rc) return rc;
    }
    // the set that is neither waiting in the queue nor (as the current set) possibly still read by the later stages of the
    // batch collected last: with one batch queued that is the other one; ev_done orders us behind those stages
    int si = -1;
    for (int k = 1; k <= kSets && si < 0; ++k) {
        const int c = (h->cur + k) % kSets;          // prefer a set other than the current one (k == kSets: the current one itself)
        bool queued = false;
        for (int q = 0; q < h->n_fifo; ++q) queued |= h->fifo[q] == c;
        if (!queued) si = c;
    }
    fqb_handle::BatchSet &B = h->sets[si];
    cudaStream_t ast = h->align_stream[si];
    CU_CHECK(cudaStreamWaitEvent(ast, B.ev_done, 0));
    // the upload runs on the copy stream (it only needs the staging arrays, free once the previous occupant's prep_kernel has
    // run), so it overlaps the search of the batch before; the align stream picks it up through ev_in
    int rc = load_set(h, si, h->copy_stream, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, on_device, packed_stride);
    if (rc) return rc;
    CU_CHECK(cudaEventRecord(B.ev_in, h->copy_stream));
    CU_CHECK(cudaStreamWaitEvent(ast, B.ev_in, 0));
    static const bool tl_on = getenv("FQB_TIMELINE") != nullptr;
    auto tl_mark = [&](cudaStream_t s_) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s_); h->tl_ev.push_back(e); };
    if (tl_on) { if (!h->tl_base) { cudaEventCreate(&h->tl_base); cudaEventRecord(h->tl_base, ast); } tl_mark(ast); }
    if ((rc = enqueue_align(h, si, ast))) return rc;
    CU_CHECK(cudaEventRecord(B.ev_align, ast));
    if (tl_on) tl_mark(ast);
    h->fifo[h->n_fifo++] = si;
    return FQB_OK;
}
int fqb_submit_pairs(fqb_handle *h, int32_t n_pairs, int32_t stride, const uint8_t *bases1, const uint8_t *quals1,
                     const int32_t *lens1, const uint8_t *bases2, const uint8_t *quals2, const int32_t *lens2, int on_device) {
    return submit_impl(h, n_pairs, stride, bases1, quals1, lens1, bases2, quals2, lens2, on_device, 0);
}
int fqb_submit_pairs_packed(fqb_handle *h, int32_t n_pairs, int32_t stride, int32_t packed_stride, const uint8_t *packed1, const uint8_t *quals1,
                            const int32_t *lens1, const uint8_t *packed2, const uint8_t *quals2, const int32_t *lens2, int on_device) {
    if (packed_stride <= 0) { set_error("fqb_submit_pairs_packed: packed_stride must be positive"); return FQB_ERR_ARG; }
    return submit_impl(h, n_pairs, stride, packed1, quals1, lens1, packed2, quals2, lens2, on_device, packed_stride);
}
int fqb_collect_pairs(fqb_handle *h, fqb_read_t *rows1, fqb_read_t *rows2) {
    if (!h || !h->n_fifo) { set_error("fqb_collect_pairs: no batch submitted"); return FQB_ERR_STATE; }
    CU_CHECK(cudaSetDevice(h->device));
    const int si = h->fifo[0];
    h->fifo[0] = h->fifo[1]; --h->n_fifo;
    use_set(h, si);
    CU_CHECK(cudaStreamWaitEvent(h->stream, h->sets[si].ev_align, 0));
    int rc = enqueue_pair(h, h->stream);
    if (!rc) rc = enqueue_sw_refine(h, h->stream);
    if (rc) return rc;
    h->batch_ready = h->align_done = h->pair_done = h->dp_done = true;
    // the rows are those fqb_align_pairs returns: before StatCollector::AddAlignment's bridge check demotes reads in place
    if (rows1 && rows2) { if ((rc = fqb_stage_fetch_rows_async(h, rows1, rows2))) return rc; }
    if (h->stats_open) rc = enqueue_stats(h, h->stream);
    if (!rc) rc = enqueue_status(h, h->stream);
    if (rc) return rc;
    h->stats_done = h->stats_open;
    CU_CHECK(cudaEventRecord(h->sets[si].ev_done, h->stream));
    return FQB_OK;
}

// pinned host memory for the FASTQ feeder's batches
void *fqb_host_alloc(size_t bytes) { void *p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
void fqb_host_free(void *p) { if (p) cudaFreeHost(p); }

uint64_t fqb_launch_count(const fqb_handle *h) { return h ? h->n_launches : 0; }

// device time (CUDA events on the handle's stream) spent in the rank-query kernels -- bwt_cal_width + queue ordering +
// bwt_match_gap fast pass -- since creation, and the number of batches it covers; the roofline's "dominant kernel" clock
int fqb_rank_query_time(const fqb_handle *hc, double *ms, uint64_t *launches) {
    fqb_handle *h = const_cast<fqb_handle *>(hc);
    if (!h) { set_error("null handle"); return FQB_ERR_ARG; }
    cudaSetDevice(h->device);
    for (auto &B : h->sets) harvest_rq(h, B, true);       // brackets still in flight are waited for and counted
    if (ms) *ms = h->rq_ms;
    if (launches) *launches = h->rq_launches;
    return FQB_OK;
}

void *fqb_stream(fqb_handle *h) { return h ? (void *)h->stream : nullptr; }

}  // extern "C"
