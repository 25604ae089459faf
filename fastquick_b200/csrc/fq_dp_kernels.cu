// CUDA kernels of the dynamic-programming rows (a10 mate rescue, a11 gapped refinement), sm_100a.
// One alignment per lane; DP rows and trace-back live in warp-interleaved global scratch so the 32
// lanes of a warp (walking their matrices in lock step) issue coalesced accesses.  Work comes from
// compacted lists through a warp-aggregated queue, as in the search kernel.
#include "fq_dp_kernels.cuh"

namespace fqb {

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ DpScratch lane_scratch(const DpPool &pool) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    DpScratch sc;
    sc.ints = pool.ints + warp * (size_t)pool.ints_per_lane * 32 + lane;
    sc.bytes = pool.bytes + warp * (size_t)pool.bytes_per_lane * 32 + lane;
    sc.n_ints = pool.ints_per_lane; sc.n_bytes = pool.bytes_per_lane; sc.istride = sc.bstride = 32;
    return sc;
}
// fast variant: DP rows in shared memory (element e of thread t at e * blockDim + t: conflict-free), trace-back bytes in
// warp-interleaved global scratch (written once, read only by the back-trace)
__device__ __forceinline__ DpScratch lane_scratch_smem(const DpPool &pool, int32_t *smem, int ints_per_lane) {
    DpScratch sc = lane_scratch(pool);
    sc.ints = smem + threadIdx.x; sc.n_ints = ints_per_lane; sc.istride = blockDim.x;
    return sc;
}

// warp-aggregated fetch of the next work item; returns false when the list is exhausted
__device__ __forceinline__ bool next_item(uint32_t *cursor, uint32_t n, uint32_t &idx) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    idx = base + __popc(m & ((1u << lane) - 1u));
    return idx < n;
}

// head of bwa_paired_sw's per-pair loop: un-filter rescued mates, collect the pairs that qualify for mate rescue
__global__ void sw_classify_kernel(DpView v, uint32_t *list, uint32_t *n_list) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false;
    if (p < (uint32_t)v.n_reads / 2) {
        PairParams pp; pp.sw_on = 1;
        fqb_read_t *p0 = v.rows + 2 * p, *p1 = p0 + 1;
        const uint8_t f0 = p0->filtered, f1 = p1->filtered;
        fqb_read_t a = *p0, b = *p1;
        need = sw_candidate(&a, &b, pp);
        if (a.filtered != f0) p0->filtered = a.filtered;
        if (b.filtered != f1) p1->filtered = b.filtered;
    }
    const unsigned m = __ballot_sync(FULL_MASK, need);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(n_list, (uint32_t)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (need) list[base + __popc(m & ((1u << lane) - 1u))] = p;
}

// kSmem: rows in shared memory, items that do not fit go to `retry` (processed by the global-scratch variant)
template <bool kSmem>
__global__ void __launch_bounds__(kDpThreads) sw_kernel(DpView v, SwParams sp, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                         uint32_t *cursor, uint32_t *err, int smem_ints, uint32_t *retry, uint32_t *n_retry) {
    extern __shared__ int32_t dp_smem[];
    DpScratch sc = kSmem ? lane_scratch_smem(pool, dp_smem, smem_ints) : lane_scratch(pool);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item(cursor, n, j)) {
        const uint32_t p = list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        bool ok = paired_sw_one(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, sc);
        if (!ok) {
            if (kSmem) retry[atomicAdd(n_retry, 1u)] = p; else atomicExch(err, p + 1);
            continue;
        }
        v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1;
    }
}

__global__ void refine_classify_kernel(DpView v, uint32_t *list, uint32_t *n_list) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false;
    if (r < (uint32_t)v.n_reads) {
        const fqb_read_t &s = v.rows[r];
        need = !s.filtered && !(s.type == kTypeNoMatch || s.type == kTypeMateSW || s.n_gapo == 0);
    }
    const unsigned m = __ballot_sync(FULL_MASK, need);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(n_list, (uint32_t)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (need) list[base + __popc(m & ((1u << lane) - 1u))] = r;
}

template <bool kSmem>
__global__ void __launch_bounds__(kDpThreads) refine_kernel(DpView v, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                             uint32_t *cursor, uint32_t *err, int smem_ints, uint32_t *retry, uint32_t *n_retry) {
    extern __shared__ int32_t dp_smem[];
    DpScratch sc = kSmem ? lane_scratch_smem(pool, dp_smem, smem_ints) : lane_scratch(pool);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item(cursor, n, j)) {
        const uint32_t r = list[j];
        fqb_read_t s = v.rows[r];
        ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = s.len; Q.strand = s.strand;
        int nc = refine_gapped(v.l_pac, v.pac, Q, &s.pos, (s.strand ? 1 : -1) * (s.n_gapo + s.n_gape), s.cigar, FQB_MAX_CIGAR, sc);
        if (nc < 0) {
            if (kSmem) retry[atomicAdd(n_retry, 1u)] = r; else atomicExch(err, r + 1);
            continue;
        }
        s.n_cigar = (uint8_t)nc; s.has_cigar = 1;
        v.rows[r] = s;
    }
}

// NM (bwa_cal_md1's count) + bwa_correct_trimmed for every read
__global__ void finish_kernel(DpView v) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (uint32_t)v.n_reads) return;
    fqb_read_t s = v.rows[r];
    if (s.type != kTypeNoMatch) {
        ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = s.len; Q.strand = s.strand;
        s.nm = (uint16_t)cal_nm(s, Q, v.l_pac, v.pac);
    }
    correct_trimmed(s);
    v.rows[r] = s;
}

// ctr: [0] n_list [1] cursor [2] n_retry [3] retry cursor (device words)
void launch_sw(const DpView &v, const SwParams &sp, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, cudaStream_t s) {
    sw_classify_kernel<<<(v.n_reads / 2 + 255) / 256, 256, 0, s>>>(v, list, ctr);
    const size_t smem = (size_t)kSwSmemInts * kDpThreads * 4;
    cudaFuncSetAttribute(sw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sw_kernel<true><<<pool.n_blocks, kDpThreads, smem, s>>>(v, sp, pool, list, ctr, ctr + 1, err, kSwSmemInts, retry, ctr + 2);
    sw_kernel<false><<<pool.n_blocks, kDpThreads, 0, s>>>(v, sp, pool, retry, ctr + 2, ctr + 3, err, 0, nullptr, nullptr);
}
void launch_refine(const DpView &v, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, int max_read_len, cudaStream_t s) {
    refine_classify_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v, list, ctr);
    const int ints = 6 * (max_read_len + 16 + 1);
    int threads = (int)((220u * 1024u) / ((size_t)ints * 4)) / 32 * 32;          // as many lanes as fit 220 KB of shared memory
    if (threads > kDpThreads) threads = kDpThreads;
    if (threads < 32) threads = 32;
    const size_t smem = (size_t)ints * threads * 4;
    cudaFuncSetAttribute(refine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    refine_kernel<true><<<pool.n_blocks, threads, smem, s>>>(v, pool, list, ctr, ctr + 1, err, ints, retry, ctr + 2);
    refine_kernel<false><<<pool.n_blocks, kDpThreads, 0, s>>>(v, pool, retry, ctr + 2, ctr + 3, err, 0, nullptr, nullptr);
    finish_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v);
}

}  // namespace fqb
