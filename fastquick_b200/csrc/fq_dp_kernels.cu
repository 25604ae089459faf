// CUDA kernels of the dynamic-programming rows (a10 mate rescue, a11 gapped refinement), sm_100a.
// Main path: one alignment per WARP (fq_dp_warp.cuh), DP rows in shared memory, trace-back in a warp-private
// slab of the global pool.  Alignments whose window does not fit go to a retry list handled by the
// one-alignment-per-lane kernels (rows and trace-back in warp-interleaved global scratch).
#include <cstdlib>
#include "fq_dp_kernels.cuh"
#include "fq_dp_warp.cuh"

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

constexpr int kWarpsPerBlock = 8;
constexpr size_t kWarpSlab = 256u * 1024u;       // per-warp global slab: path ops + trace-back matrix

// shared memory of a block: kWarpsPerBlock x smem_ints words of DP rows, then kWarpsPerBlock x ref_cap bytes of reference codes
__device__ __forceinline__ WarpDp warp_scratch(const DpPool &pool, int32_t *smem, int smem_ints, int ref_cap) {
    const int wid = threadIdx.x >> 5;
    WarpDp w;
    w.sm = smem + (size_t)wid * smem_ints; w.n_ints = smem_ints;
    w.refc = reinterpret_cast<uint8_t *>(smem + (size_t)kWarpsPerBlock * smem_ints) + (size_t)wid * ref_cap; w.n_refc = ref_cap;
    w.gb = pool.bytes + ((size_t)blockIdx.x * kWarpsPerBlock + wid) * kWarpSlab; w.n_bytes = (int)kWarpSlab;
    w.lane = threadIdx.x & 31;
    return w;
}
__device__ __forceinline__ bool next_item_warp(uint32_t *cursor, uint32_t n, uint32_t &idx) {
    uint32_t j = 0;
    if ((threadIdx.x & 31) == 0) j = atomicAdd(cursor, 1u);
    idx = __shfl_sync(FULL_MASK, j, 0);
    return idx < n;
}

// mate rescue, one pair per warp; pairs whose window exceeds the shared-memory rows go to `retry`
__global__ void __launch_bounds__(kWarpsPerBlock * 32) sw_warp_kernel(DpView v, SwParams sp, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                                       uint32_t *cursor, int smem_ints, int ref_cap, uint32_t *retry, uint32_t *n_retry) {
    extern __shared__ int32_t dp_smem[];
    const WarpDp w = warp_scratch(pool, dp_smem, smem_ints, ref_cap);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item_warp(cursor, n, j)) {
        const uint32_t p = list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        WarpSwCore core{w};
        const bool ok = paired_sw_pair(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, core);
        if (w.lane == 0) {
            if (!ok) retry[atomicAdd(n_retry, 1u)] = p;
            else { v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1; }
        }
        __syncwarp();
    }
}

// one pair per lane, everything in global scratch (retry list)
__global__ void __launch_bounds__(kDpThreads) sw_kernel(DpView v, SwParams sp, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                         uint32_t *cursor, uint32_t *err) {
    DpScratch sc = lane_scratch(pool);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item(cursor, n, j)) {
        const uint32_t p = list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        bool ok = paired_sw_one(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, sc);
        if (!ok) { atomicExch(err, p + 1); continue; }
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

__global__ void __launch_bounds__(kWarpsPerBlock * 32) refine_warp_kernel(DpView v, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                                           uint32_t *cursor, int smem_ints, int ref_cap, uint32_t *retry, uint32_t *n_retry) {
    extern __shared__ int32_t dp_smem[];
    const WarpDp w = warp_scratch(pool, dp_smem, smem_ints, ref_cap);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item_warp(cursor, n, j)) {
        const uint32_t r = list[j];
        fqb_read_t s = v.rows[r];
        ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = s.len; Q.strand = s.strand;
        const int nc = warp_refine_gapped(v.l_pac, v.pac, Q, &s.pos, (s.strand ? 1 : -1) * (s.n_gapo + s.n_gape), s.cigar, FQB_MAX_CIGAR, w);
        if (w.lane == 0) {
            if (nc < 0) retry[atomicAdd(n_retry, 1u)] = r;
            else { s.n_cigar = (uint8_t)nc; s.has_cigar = 1; v.rows[r] = s; }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kDpThreads) refine_kernel(DpView v, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                             uint32_t *cursor, uint32_t *err) {
    DpScratch sc = lane_scratch(pool);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item(cursor, n, j)) {
        const uint32_t r = list[j];
        fqb_read_t s = v.rows[r];
        ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = s.len; Q.strand = s.strand;
        int nc = refine_gapped(v.l_pac, v.pac, Q, &s.pos, (s.strand ? 1 : -1) * (s.n_gapo + s.n_gape), s.cigar, FQB_MAX_CIGAR, sc);
        if (nc < 0) { atomicExch(err, r + 1); continue; }
        s.n_cigar = (uint8_t)nc; s.has_cigar = 1;
        v.rows[r] = s;
    }
}

// ---- XA: the other hits of reads that keep a multi list (rare: n_occ <= n_multi + 1) ----
__global__ void multi_classify_kernel(DpView v, uint32_t *list, uint32_t *ctr) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (uint32_t)v.n_reads) return;
    const fqb_read_t &s = v.rows[r];
    if (s.filtered || s.type == kTypeNoMatch || s.n_multi == 0) return;
    const uint32_t i = atomicAdd(ctr, 1u);
    list[2 * i] = r;
    list[2 * i + 1] = atomicAdd(ctr + 1, (uint32_t)s.n_multi);
}
__global__ void __launch_bounds__(kDpThreads) multi_kernel(DpView v, MultiView mv, DpPool pool, const uint32_t *list, uint32_t *ctr, MultiOut *out,
                                                            uint32_t out_cap, uint32_t *err) {
    DpScratch sc = lane_scratch(pool);
    const uint32_t n = ctr[0];
    uint32_t it;
    while (next_item(ctr + 2, n, it)) {
        const uint32_t r = list[2 * it], base = list[2 * it + 1];
        const fqb_read_t s = v.rows[r];
        const int spill = mv.spill_slot[r];
        const Hit *al = spill >= 0 ? mv.aln_big + (size_t)spill * mv.aln_big_cap : mv.aln + (size_t)r * mv.aln_cap;
        const int na = mv.n_aln[r], len = s.clip_len;          // positions and CIGARs use the trimmed read, as the reference does at that point
        int z = 0;
        for (int k = 0; k < na && z < s.n_multi; ++k)
            for (uint32_t row = al[k].k; row <= al[k].l && z < s.n_multi; ++row) {
                if (row == s.sa) continue;
                if (base + z >= out_cap) { atomicExch(err, r + 1); return; }
                MultiOut o;
                o.read = r; o.j = (uint8_t)z; o.strand = al[k].a; o.gap = (uint8_t)(al[k].n_gapo + al[k].n_gape); o.mm = al[k].n_mm;
                o.pos = hit_position(mv.bwt, al[k].a, row, len);
                o.n_cigar = 0; o.has_cigar = 0; o.pad_[0] = o.pad_[1] = 0;
                for (int c = 0; c < FQB_MAX_CIGAR; ++c) o.cigar[c] = 0;
                if (o.gap) {
                    ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = len; Q.strand = o.strand;
                    const int nc = refine_gapped(v.l_pac, v.pac, Q, &o.pos, (o.strand ? 1 : -1) * (int)o.gap, o.cigar, FQB_MAX_CIGAR, sc);
                    if (nc < 0) { atomicExch(err, r + 1); return; }
                    o.n_cigar = (uint8_t)nc; o.has_cigar = 1;
                }
                out[base + z] = o;
                ++z;
            }
    }
}
void launch_multi(const DpView &v, const MultiView &mv, const DpPool &pool, uint32_t *list, uint32_t *ctr, MultiOut *out, uint32_t out_cap,
                  uint32_t *err, cudaStream_t s) {
    cudaMemsetAsync(ctr, 0, 3 * 4, s);
    multi_classify_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v, list, ctr);
    multi_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, mv, pool, list, ctr, out, out_cap, err);
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

// as many warp slabs as the byte pool holds, at most four blocks per SM (pool.n_blocks = 2 per SM)
static int warp_blocks(const DpPool &pool) {
    const size_t pool_bytes = (size_t)pool.n_blocks * kDpThreads * pool.bytes_per_lane;
    size_t nb = pool_bytes / (kWarpSlab * kWarpsPerBlock);
    if (nb > (size_t)pool.n_blocks * 2) nb = (size_t)pool.n_blocks * 2;
    return (int)(nb < 1 ? 1 : nb);
}

// ctr: [0] n_list [1] cursor [2] n_retry [3] retry cursor (device words)
void launch_sw(const DpView &v, const SwParams &sp, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, cudaStream_t s) {
    sw_classify_kernel<<<(v.n_reads / 2 + 255) / 256, 256, 0, s>>>(v, list, ctr);
    const int ints = 2 * kSwSmemInts;                                    // H and E rows of a <= 702-column window
    const int ref_cap = kSwSmemInts;                                     // reference codes of the window, one byte each
    const size_t smem = (size_t)ints * kWarpsPerBlock * 4 + (size_t)ref_cap * kWarpsPerBlock;
    cudaFuncSetAttribute(sw_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (getenv("FQB_DP_NO_WARP"))        // debugging aid: everything through the one-alignment-per-lane kernels
        sw_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, sp, pool, list, ctr, ctr + 1, err);
    else {
        sw_warp_kernel<<<warp_blocks(pool), kWarpsPerBlock * 32, smem, s>>>(v, sp, pool, list, ctr, ctr + 1, ints, ref_cap, retry, ctr + 2);
        sw_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, sp, pool, retry, ctr + 2, ctr + 3, err);
    }
}
void launch_refine(const DpView &v, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, int max_read_len, cudaStream_t s) {
    refine_classify_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v, list, ctr);
    int ints = 6 * (max_read_len + 16 + 1);                              // M/I/D rows, current and previous
    if ((size_t)ints * kWarpsPerBlock * 4 > 200u * 1024u) ints = (int)(200u * 1024u / (kWarpsPerBlock * 4));
    const int ref_cap = (ints / 6 + 3) & ~3;                             // window columns + 1, padded to a word
    const size_t smem = (size_t)ints * kWarpsPerBlock * 4 + (size_t)ref_cap * kWarpsPerBlock;
    cudaFuncSetAttribute(refine_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (getenv("FQB_DP_NO_WARP"))
        refine_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, pool, list, ctr, ctr + 1, err);
    else {
        refine_warp_kernel<<<warp_blocks(pool), kWarpsPerBlock * 32, smem, s>>>(v, pool, list, ctr, ctr + 1, ints, ref_cap, retry, ctr + 2);
        refine_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, pool, retry, ctr + 2, ctr + 3, err);
    }
    finish_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v);
}

}  // namespace fqb
