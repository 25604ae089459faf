// CUDA kernels of the paired-end resolution stage (rows a6-a9), sm_100a.
//   se_count / scan / se_multi_seq / se_final : bwa_aln2seq_core with the GLOBAL drand48 stream, made
//        parallel: almost every read consumes exactly two draws, so offsets are a prefix sum; the rare
//        reads with several best intervals are resolved by one thread walking them in read order
//        (closed-form LCG jump-ahead); every read then jumps to its own offset.
//   isize_hist  : insert sizes of confidently mapped pairs -> histogram (host runs infer_isize's libm part)
//   pair_kernel : pairing() per pair + multi-hit counts
#include "fq_pair_kernels.cuh"

namespace fqb {

__device__ __forceinline__ const Hit *hits_of(const PeView &v, uint32_t r) {
    int s = v.spill_slot[r];
    return s >= 0 ? v.aln_big + (size_t)s * v.aln_big_cap : v.aln + (size_t)r * v.aln_cap;
}
__device__ __forceinline__ int n_hits_of(const PeView &v, uint32_t r) { return v.filtered[r] ? 0 : v.n_aln[r]; }

// packed[r] = provisional draw count (low 32) | "needs sequential resolution" flag (high 32)
__global__ void se_count_kernel(PeView v, uint64_t *packed) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (uint32_t)v.n_reads) return;
    int na = n_hits_of(v, r);
    uint64_t out = 0;
    if (na > 0) {
        int nb = count_best(hits_of(v, r), na);
        out = nb == 1 ? 2ull : (1ull << 32);
    }
    packed[r] = out;
}

// ---- exclusive scan of u64 (both halves scan independently as long as the low half stays < 2^32)
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) scan_blocks_kernel(const uint64_t *in, uint64_t *out, uint64_t *block_sums, int n) {
    __shared__ uint64_t warp_tot[32];
    const int i = blockIdx.x * kScanBlock + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t v = i < n ? in[i] : 0, x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        uint64_t t = warp_tot[lane], y = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t u = __shfl_up_sync(0xffffffffu, y, d); if (lane >= d) y += u; }
        warp_tot[lane] = y - t;
        if (lane == 31 && block_sums) block_sums[blockIdx.x] = y;
    }
    __syncthreads();
    if (i < n) out[i] = x - v + warp_tot[w];
}
__global__ void scan_add_kernel(uint64_t *out, const uint64_t *block_offs, int n) {
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += block_offs[blockIdx.x];
}
void exclusive_scan_u64(const uint64_t *in, uint64_t *out, uint64_t *tmp /* 2 * ceil(n/1024) + 2 */, int n, cudaStream_t s) {
    int nb = (n + kScanBlock - 1) / kScanBlock;
    if (nb < 1) nb = 1;
    scan_blocks_kernel<<<nb, kScanBlock, 0, s>>>(in, out, tmp, n);
    if (nb > 1) {
        if (nb > kScanBlock) return;   // n <= 1M entries per call by construction (checked by the caller)
        scan_blocks_kernel<<<1, kScanBlock, 0, s>>>(tmp, tmp + nb, nullptr, nb);
        scan_add_kernel<<<nb, kScanBlock, 0, s>>>(out, tmp + nb, n);
    }
}

__global__ void se_multi_list_kernel(PeView v, const uint64_t *packed, const uint64_t *scanned, uint32_t *list) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (uint32_t)v.n_reads) return;
    if (packed[r] >> 32) list[scanned[r] >> 32] = r;
}

// one thread: reads with several best-score intervals, in read order
__global__ void se_multi_seq_kernel(PeView v, const uint64_t *packed, const uint64_t *scanned, const uint32_t *list,
                                    uint64_t *cum_extra, RngState rng_, const uint64_t *calls, uint64_t *totals) {
    const RngState rng{rng_.x0, rng_.calls + (calls ? *calls : 0)};
    if (blockIdx.x || threadIdx.x) return;
    const int n = v.n_reads;
    const uint32_t n_multi = n ? (uint32_t)((scanned[n - 1] + packed[n - 1]) >> 32) : 0;
    uint64_t extra = 0;
    for (uint32_t m = 0; m < n_multi; ++m) {
        const uint32_t r = list[m];
        const uint64_t off = rng.calls + (scanned[r] & 0xffffffffull) + extra;
        fqb_read_t tmp;
        tmp.sa = 0; tmp.c1 = tmp.c2 = 0;
        extra += se_choose(hits_of(v, r), n_hits_of(v, r), lcg_advance(rng.x0, off), tmp);
        cum_extra[m] = extra;
    }
    totals[0] = n ? ((scanned[n - 1] + packed[n - 1]) & 0xffffffffull) + extra : 0;   // draws consumed by this batch
    totals[1] = n_multi;
}

// SE pass of bwa_cal_pac_pos_pe (src/BwtMapper.cpp:744-776) for read r
__global__ void se_final_kernel(PeView v, SeParams sp, const uint64_t *packed, const uint64_t *scanned, const uint32_t *list,
                                const uint64_t *cum_extra, const uint64_t *totals, RngState rng_, const uint64_t *calls, uint32_t *err_flag) {
    const RngState rng{rng_.x0, rng_.calls + (calls ? *calls : 0)};
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (uint32_t)v.n_reads) return;
    fqb_read_t row;
    uint32_t *rw = reinterpret_cast<uint32_t *>(&row);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(row) / 4); ++i) rw[i] = 0;
    row.len = v.len[r]; row.full_len = v.full_len[r]; row.clip_len = row.len;
    row.filtered = v.filtered[r];
    row.extra_flag = v.single_end ? 0 : (kSamPaired | ((r & 1) ? kSamRead2 : kSamRead1));
    const int na = n_hits_of(v, r);
    row.n_aln = (uint16_t)(na > 65535 ? 65535 : na);
    if (na > 0) {
        const uint32_t n_multi = (uint32_t)totals[1];
        uint32_t lo = 0, hi = n_multi;                 // multi reads strictly before r
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (list[mid] < r) lo = mid + 1; else hi = mid; }
        const uint64_t extra = lo ? cum_extra[lo - 1] : 0;
        const uint64_t off = rng.calls + (scanned[r] & 0xffffffffull) + extra;
        const uint32_t used = se_choose(hits_of(v, r), na, lcg_advance(rng.x0, off), row);
        if (!(packed[r] >> 32) && used != 2) atomicExch(err_flag, r + 1);   // first draw was exactly 0.0 (p = 2^-48): host reroutes
        const int max_diff = sp.maxdiff[row.len];
        row.pos = hit_position(sp.bwt, row.strand, row.sa, row.len);
        row.seQ = row.mapQ = (uint8_t)approx_mapq(row.c1, row.c2, row.n_mm, max_diff, sp.g_log_n);
        if (v.single_end) {
            // bwa_aln2seq_core(..., set_main = 1, n_multi = N_OCC = 3): keep the other hits only when there are at most N_OCC of them
            const Hit *al = hits_of(v, r);
            uint32_t n_occ = 0, others = 0;
            for (int k = 0; k < na; ++k) {
                n_occ += al[k].l - al[k].k + 1;
                others += al[k].l - al[k].k + 1 - ((row.sa >= al[k].k && row.sa <= al[k].l) ? 1u : 0u);
            }
            row.n_multi = n_occ > 3u + 1u ? 0 : (uint8_t)(others < 3u ? others : 3u);
        }
    }
    v.rows[r] = row;
}

// infer_isize's collection loop (libbwa/bwape.c:58-67)
__global__ void isize_hist_kernel(PeView v, uint32_t *hist, uint32_t *max_len) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ml = 1;
    if (p < (uint32_t)v.n_reads / 2) {
        const fqb_read_t &a = v.rows[2 * p], &b = v.rows[2 * p + 1];
        if (a.mapQ >= 20 && b.mapQ >= 20) {
            uint64_t x = a.pos < b.pos ? (uint64_t)b.pos + b.len - a.pos : (uint64_t)a.pos + a.len - b.pos;
            if (x < 100000) atomicAdd(hist + x, 1u);
        }
        ml = (uint32_t)(a.len > b.len ? a.len : b.len);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { uint32_t o = __shfl_xor_sync(0xffffffffu, ml, d); ml = o > ml ? o : ml; }
    if ((threadIdx.x & 31) == 0) atomicMax(max_len, ml);
}

__global__ void __launch_bounds__(128) pair_kernel(PeView v, DevBwt b0, DevBwt b1, const PairParams *ppp, uint32_t *big_list, uint32_t *n_big,
                                                    uint32_t *sw_list, uint32_t *n_sw) {
    const PairParams pp = *ppp;
    __shared__ DevBwt s_bwt[2];
    if (threadIdx.x == 0) { s_bwt[0] = b0; s_bwt[1] = b1; }
    __syncthreads();
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (uint32_t)v.n_reads / 2) return;
    uint64_t arr[kPairArrCap];
    fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
    bool ok = pair_one(s_bwt, &r0, &r1, hits_of(v, 2 * p), n_hits_of(v, 2 * p), hits_of(v, 2 * p + 1), n_hits_of(v, 2 * p + 1),
                       pp, arr, kPairArrCap);
    if (!ok) {                      // many hit positions (repeats): pair_big_kernel takes it; n_big[1] flags more than it can hold
        const uint32_t slot = atomicAdd(n_big, 1u);
        if (slot < kPairBigMax) big_list[slot] = p; else atomicExch(n_big + 1, p + 1);
        return;
    }
    if (sw_candidate(&r0, &r1, pp)) sw_list[atomicAdd(n_sw, 1u)] = p;
    v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1;
}

// pairs with more hit positions than a thread sorts in registers/local memory: one thread each, scratch in global memory
__global__ void pair_big_kernel(PeView v, DevBwt b0, DevBwt b1, const PairParams *ppp, const uint32_t *big_list, const uint32_t *n_big,
                                uint64_t *scratch, size_t scratch_per_pair, uint32_t *sw_list, uint32_t *n_sw) {
    const PairParams pp = *ppp;
    __shared__ DevBwt s_bwt[2];
    if (threadIdx.x == 0) { s_bwt[0] = b0; s_bwt[1] = b1; }
    __syncthreads();
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *n_big || j >= kPairBigMax) return;
    uint32_t p = big_list[j];
    fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
    pair_one(s_bwt, &r0, &r1, hits_of(v, 2 * p), n_hits_of(v, 2 * p), hits_of(v, 2 * p + 1), n_hits_of(v, 2 * p + 1), pp,
             scratch + (size_t)j * scratch_per_pair, (int)scratch_per_pair);
    if (sw_candidate(&r0, &r1, pp)) sw_list[atomicAdd(n_sw, 1u)] = p;
    v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1;
}

// the part of the SE pass that does not depend on the stream position: provisional draw counts, their prefix sums, the list
// of reads with several best intervals (in a sharded run it overlaps the wait for the previous batch's owner)
void launch_se_prepare(const PeView &v, PeScratch &sc, cudaStream_t s) {
    const int n = v.n_reads, tb = 256, nb = (n + tb - 1) / tb;
    se_count_kernel<<<nb, tb, 0, s>>>(v, sc.packed);
    exclusive_scan_u64(sc.packed, sc.scanned, sc.scan_tmp, n, s);
    se_multi_list_kernel<<<nb, tb, 0, s>>>(v, sc.packed, sc.scanned, sc.multi_list);
}
void launch_se_finish(const PeView &v, const SeParams &sp, const RngState &rng, const uint64_t *calls, PeScratch &sc, cudaStream_t s) {
    const int n = v.n_reads;
    se_multi_seq_kernel<<<1, 32, 0, s>>>(v, sc.packed, sc.scanned, sc.multi_list, sc.cum_extra, rng, calls, sc.totals);
    se_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(v, sp, sc.packed, sc.scanned, sc.multi_list, sc.cum_extra, sc.totals, rng, calls, sc.err_flag);
}
void launch_isize_hist(const PeView &v, uint32_t *hist, uint32_t *max_len, cudaStream_t s) {
    const int np = v.n_reads / 2;
    isize_hist_kernel<<<(np + 255) / 256, 256, 0, s>>>(v, hist, max_len);
}
void launch_pair(const PeView &v, const DevBwt bwt[2], const PairParams *pp, uint32_t *big_list, uint32_t *n_big, uint32_t *sw_list,
                 uint32_t *n_sw, cudaStream_t s) {
    const int np = v.n_reads / 2;
    pair_kernel<<<(np + 127) / 128, 128, 0, s>>>(v, bwt[0], bwt[1], pp, big_list, n_big, sw_list, n_sw);
}
void launch_pair_big(const PeView &v, const DevBwt bwt[2], const PairParams *pp, const uint32_t *big_list, const uint32_t *n_big,
                     uint64_t *scratch, size_t scratch_per_pair, uint32_t *sw_list, uint32_t *n_sw, cudaStream_t s) {
    pair_big_kernel<<<(kPairBigMax + 63) / 64, 64, 0, s>>>(v, bwt[0], bwt[1], pp, big_list, n_big, scratch, scratch_per_pair, sw_list, n_sw);
}

}  // namespace fqb
