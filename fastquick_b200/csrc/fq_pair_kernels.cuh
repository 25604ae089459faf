// Launchers of the paired-end resolution stage (fq_pair_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "fq_device_pair.cuh"

namespace fqb {

struct PeView {
    int n_reads;
    const Hit *aln; int aln_cap;
    const Hit *aln_big; int aln_big_cap;
    const int32_t *spill_slot;      // per read: row of aln_big, or -1
    const int32_t *n_aln;
    const uint8_t *filtered;
    const int32_t *len, *full_len;
    fqb_read_t *rows;
    int single_end;                 // SingleEndMapper: no pairing flags, multi list of up to N_OCC hits (src/BwtMapper.cpp:1340)
};
struct SeParams {
    DevBwt bwt[2];
    const int32_t *maxdiff;         // per read length
    const int32_t *g_log_n;
};
struct RngState { uint64_t x0; uint64_t calls; };   // srand48 state and draws consumed by earlier batches
struct PeScratch {
    uint64_t *packed, *scanned, *scan_tmp, *cum_extra, *totals;
    uint32_t *multi_list, *err_flag;
};

// calls: optional DEVICE word added to rng.calls (the stream position handed from batch to batch stays on the device)
void launch_se_prepare(const PeView &v, PeScratch &sc, cudaStream_t s);
void launch_se_finish(const PeView &v, const SeParams &sp, const RngState &rng, const uint64_t *calls, PeScratch &sc, cudaStream_t s);
void launch_isize_hist(const PeView &v, uint32_t *hist, uint32_t *max_len, cudaStream_t s);
// pp: DEVICE pointer (written by the pair stage's host callback once infer_isize has run).  n_big: [0] pairs deferred to
// pair_big_kernel, [1] != 0 when more than kPairBigMax pairs asked for it (a limit error).  launch_pair_big always runs its
// fixed grid; it reads the count on the device.
constexpr uint32_t kPairBigMax = 4096;
void launch_pair(const PeView &v, const DevBwt bwt[2], const PairParams *pp, uint32_t *big_list, uint32_t *n_big, uint32_t *sw_list,
                 uint32_t *n_sw, cudaStream_t s);
void launch_pair_big(const PeView &v, const DevBwt bwt[2], const PairParams *pp, const uint32_t *big_list, const uint32_t *n_big,
                     uint64_t *scratch, size_t scratch_per_pair, uint32_t *sw_list, uint32_t *n_sw, cudaStream_t s);

}  // namespace fqb
