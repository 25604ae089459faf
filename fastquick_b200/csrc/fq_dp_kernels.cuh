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

// Alternative hits of a read for the XA tag (bwa_aln2seq_core's multi list, libbwa/bwase.c:47-95; positions as in
// bwa_cal_pac_pos_pe, src/BwtMapper.cpp:877-885; CIGARs of gapped ones as in bwa_refine_gapped, libbwa/bwase.c:361-368)
struct MultiOut {
    uint32_t read, pos;
    uint8_t strand, gap, mm, n_cigar;
    uint8_t j, has_cigar, pad_[2];
    uint16_t cigar[FQB_MAX_CIGAR];
};
struct MultiView {
    const Hit *aln; int aln_cap;
    const Hit *aln_big; int aln_big_cap;
    const int32_t *spill_slot, *n_aln;
    DevBwt bwt[2];
};
// list: 2 words per selected read (read, first output slot); ctr: [0] n_list [1] n_out [2] cursor
void launch_multi(const DpView &v, const MultiView &mv, const DpPool &pool, uint32_t *list, uint32_t *ctr, MultiOut *out, uint32_t out_cap,
                  uint32_t *err, cudaStream_t s);

// huge_mem: launch_sw_huge_bytes() bytes of device memory for the very-wide-window path
// sp: DEVICE pointer (the batch's parameters are written by the pair stage's host callback; sp->on == 0 turns the stage off)
void launch_sw(const DpView &v, const SwParams *sp, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, void *huge_mem,
               int max_read_len, cudaStream_t s);
size_t launch_sw_huge_bytes();
void launch_refine(const DpView &v, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, int max_read_len, cudaStream_t s);

}  // namespace fqb
