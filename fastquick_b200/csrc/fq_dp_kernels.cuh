// Launchers of the DP rows (fq_dp_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "fq_device_dp.cuh"

namespace fqb {

constexpr int kDpThreads = 64;
constexpr int kSwSmemInts = 704;     // mate-rescue windows up to 702 columns keep their rows in shared memory (176 KB per block)

struct DpView {
    int n_reads, lpad;
    const uint8_t *codes;     // nt4, read orientation
    const uint8_t *pac; int64_t l_pac;
    fqb_read_t *rows;
};
// per-lane scratch, interleaved per warp: ints_per_lane x 4 B + bytes_per_lane x 1 B for each of n_blocks*kDpThreads lanes
struct DpPool {
    int32_t *ints; uint8_t *bytes;
    int ints_per_lane, bytes_per_lane, n_blocks;
};

void launch_sw(const DpView &v, const SwParams &sp, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, cudaStream_t s);
void launch_refine(const DpView &v, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, int max_read_len, cudaStream_t s);

}  // namespace fqb
