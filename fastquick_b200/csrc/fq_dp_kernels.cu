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
__global__ void sw_classify_kernel(DpView v, const SwParams *spp, uint32_t *list, uint32_t *n_list) {
    if (!spp->on) return;              // no mate rescue for this batch (infer_isize failed, or --no-sw)
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

constexpr int kWarpsPerBlock = 4;       // small blocks (<= 16 K registers): one fits beside a search grid that leaves a fifth of an SM free
constexpr uint32_t kHugeMaxPairs = 64, kHugeMaxJobs = 128, kHugeMaxSlices = 1u << 16;   // very wide windows per batch (rare)
constexpr size_t kWarpSlab = 256u * 1024u;       // per-warp global slab: path ops + trace-back matrix

// shared memory of a block: kWarpsPerBlock x smem_ints words of DP rows / trace-back, then kWarpsPerBlock x ref_cap bytes of
// reference codes, then kWarpsPerBlock x ops_cap bytes of path ops
__device__ __forceinline__ WarpDp warp_scratch(const DpPool &pool, int32_t *smem, int smem_ints, int ref_cap, int ops_cap) {
    const int wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;        // blocks of 1 .. kWarpsPerBlock warps (fewer when a warp's rows are large)
    WarpDp w;
    w.sm = smem + (size_t)wid * smem_ints; w.n_ints = smem_ints;
    w.refc = reinterpret_cast<uint8_t *>(smem + (size_t)wpb * smem_ints) + (size_t)wid * ref_cap; w.n_refc = ref_cap;
    w.ops = reinterpret_cast<uint8_t *>(smem + (size_t)wpb * smem_ints) + (size_t)wpb * ref_cap + (size_t)wid * ops_cap; w.n_ops = ops_cap;
    w.gb = pool.bytes + ((size_t)blockIdx.x * wpb + wid) * kWarpSlab; w.n_bytes = (int)kWarpSlab;
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
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4) sw_warp_kernel(DpView v, const SwParams *spp, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                                       uint32_t *cursor, int smem_ints, int ref_cap, int ops_cap, uint32_t *retry, uint32_t *n_retry) {
    const SwParams sp = *spp;
    extern __shared__ int32_t dp_smem[];
    const WarpDp w = warp_scratch(pool, dp_smem, smem_ints, ref_cap, ops_cap);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item_warp(cursor, n, j)) {
        const uint32_t p = list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        WarpSwCore core{w, nullptr, p, 0};
        const bool ok = paired_sw_pair(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, core);
        if (w.lane == 0) {
            if (!ok) retry[atomicAdd(n_retry, 1u)] = p;
            else { v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1; }
        }
        __syncwarp();
    }
}

// one pair per lane, everything in global scratch (retry list)
__global__ void __launch_bounds__(kDpThreads) sw_kernel(DpView v, const SwParams *spp, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                         uint32_t *cursor, uint32_t *err, uint32_t *huge_list, uint32_t *n_huge) {
    const SwParams sp = *spp;
    DpScratch sc = lane_scratch(pool);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item(cursor, n, j)) {
        const uint32_t p = list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        bool ok = paired_sw_one(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, sc);
        if (!ok) {                      // a window wider than the per-lane scratch: the sliced-scan path takes it
            const uint32_t slot = atomicAdd(n_huge, 1u);
            if (slot < kHugeMaxPairs) huge_list[slot] = p; else atomicExch(err, p + 1);
            continue;
        }
        v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1;
    }
}

// ---- very wide mate-rescue windows (fq_dp_warp.cuh): prepare the scan jobs, scan the slices, finish the pairs ----
struct HugeBuf { HugeJob *jobs; uint32_t *slice_job; ScanBest *slice_best; uint32_t *ctr; };   // ctr: [0] n_jobs [1] n_slices [2] scan cursor [3] finish cursor

__global__ void sw_huge_prepare_kernel(DpView v, const SwParams *spp, const uint32_t *huge_list, const uint32_t *n_huge, HugeBuf hb, int ref_cap, uint32_t *err) {
    const SwParams sp = *spp;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *n_huge < kHugeMaxPairs ? *n_huge : kHugeMaxPairs;
    if (t >= n) return;
    const uint32_t p = huge_list[t];
    const fqb_read_t r[2] = {v.rows[2 * p], v.rows[2 * p + 1]};
    for (int k = 0; k < 2; ++k) {
        const fqb_read_t &pref = r[1 - k], &pm = r[k];
        if (pref.type == kTypeNoMatch) continue;
        int64_t b, e; int strand;
        sw_window(pref, pm, sp, b, e, strand);
        const int reglen = (int)(e - b), len = pm.len;
        if (reglen < 20 || sp.l_pac - b < len) continue;                                   // bwa_sw_core's own early exits
        const int64_t end = b + reglen < sp.l_pac ? b + reglen : sp.l_pac;
        const int wlen = (int)(end - b);
        if (wlen + 2 <= ref_cap) continue;                                                  // the finishing warp aligns it in shared memory
        if (len + 11 * len / 9 + 2 > kScanOverlap) { atomicExch(err, p + 1); return; }     // overlap argument needs short reads
        const uint32_t j = atomicAdd(hb.ctr, 1u);
        const int n_sl = (wlen + kScanCore - 1) / kScanCore;
        const uint32_t s0 = atomicAdd(hb.ctr + 1, (uint32_t)n_sl);
        if (j >= kHugeMaxJobs || s0 + n_sl > kHugeMaxSlices) { atomicExch(err, p + 1); return; }
        HugeJob J; J.pair = p; J.k = k; J.beg = b; J.reglen = reglen; J.strand = strand; J.n_slices = n_sl; J.first_slice = (int)s0;
        hb.jobs[j] = J;
        for (int q = 0; q < n_sl; ++q) hb.slice_job[s0 + q] = j;
    }
}

__global__ void __launch_bounds__(4 * 32) sw_huge_scan_kernel(DpView v, const SwParams *spp, HugeBuf hb, int smem_ints, int ref_cap) {
    const SwParams sp = *spp;
    extern __shared__ int32_t dp_smem[];
    const int wid = threadIdx.x >> 5;
    WarpDp w;
    w.sm = dp_smem + (size_t)wid * smem_ints; w.n_ints = smem_ints;
    w.refc = reinterpret_cast<uint8_t *>(dp_smem + (size_t)4 * smem_ints) + (size_t)wid * ref_cap; w.n_refc = ref_cap;
    w.ops = nullptr; w.n_ops = 0; w.gb = nullptr; w.n_bytes = 0; w.lane = threadIdx.x & 31;
    const uint32_t n_sl = hb.ctr[1] < kHugeMaxSlices ? hb.ctr[1] : kHugeMaxSlices;
    uint32_t sidx;
    while (next_item_warp(hb.ctr + 2, n_sl, sidx)) {
        const HugeJob J = hb.jobs[hb.slice_job[sidx]];
        const int q = (int)sidx - J.first_slice;
        const fqb_read_t pm = v.rows[2 * J.pair + J.k];
        ReadSeq Q; Q.fwd = v.codes + (size_t)(2 * J.pair + J.k) * v.lpad; Q.len = pm.len; Q.strand = J.strand;
        RefWin R; R.pac = v.pac; R.beg = J.beg;
        { const int64_t end = J.beg + J.reglen < sp.l_pac ? J.beg + J.reglen : sp.l_pac; R.l = (int)(end - J.beg); }
        const int lo = q * kScanCore + 1, hi = (q + 1) * kScanCore < R.l ? (q + 1) * kScanCore : R.l;
        // bwa_sw_core's N check is repeated by the finishing kernel; a window that fails it is never looked up
        const ScanBest sb = warp_local_scan_slice(R, lo, hi, Q, pm.len, w);
        if (w.lane == 0) hb.slice_best[sidx] = sb;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4) sw_huge_finish_kernel(DpView v, const SwParams *spp, DpPool pool, const uint32_t *huge_list, const uint32_t *n_huge,
                                                                              HugeBuf hb, int smem_ints, int ref_cap, int ops_cap, uint32_t *err) {
    const SwParams sp = *spp;
    extern __shared__ int32_t dp_smem[];
    const WarpDp w = warp_scratch(pool, dp_smem, smem_ints, ref_cap, ops_cap);
    const uint32_t n = *n_huge < kHugeMaxPairs ? *n_huge : kHugeMaxPairs;
    HugeView hv; hv.jobs = hb.jobs; hv.slice_best = hb.slice_best; hv.n_jobs = (int)(hb.ctr[0] < kHugeMaxJobs ? hb.ctr[0] : kHugeMaxJobs);
    uint32_t j;
    while (next_item_warp(hb.ctr + 3, n, j)) {
        const uint32_t p = huge_list[j];
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        WarpSwCore core{w, &hv, p, 0};
        const bool ok = paired_sw_pair(v.pac, &r0, &r1, v.codes + (size_t)(2 * p) * v.lpad, v.codes + (size_t)(2 * p + 1) * v.lpad, sp, core);
        if (w.lane == 0) {
            if (!ok) atomicExch(err, p + 1);
            else { v.rows[2 * p] = r0; v.rows[2 * p + 1] = r1; }
        }
        __syncwarp();
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

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 6) refine_warp_kernel(DpView v, DpPool pool, const uint32_t *list, const uint32_t *n_list,
                                                                           uint32_t *cursor, int smem_ints, int ref_cap, int ops_cap, uint32_t *retry, uint32_t *n_retry) {
    extern __shared__ int32_t dp_smem[];
    const WarpDp w = warp_scratch(pool, dp_smem, smem_ints, ref_cap, ops_cap);
    const uint32_t n = *n_list;
    uint32_t j;
    while (next_item_warp(cursor, n, j)) {
        const uint32_t r = list[j];
        fqb_read_t *row = v.rows + r;                    // every lane reads the four scalars it needs; only lane 0 holds a CIGAR
        ReadSeq Q; Q.fwd = v.codes + (size_t)r * v.lpad; Q.len = row->len; Q.strand = row->strand;
        uint32_t pos = row->pos;
        uint16_t cigar[FQB_MAX_CIGAR];
        if (w.lane == 0)                                 // slots beyond n_cigar keep what they held (the per-lane kernel rewrites the whole row)
            for (int k = 0; k < FQB_MAX_CIGAR; ++k) cigar[k] = row->cigar[k];
        const int nc = warp_refine_gapped(v.l_pac, v.pac, Q, &pos, (Q.strand ? 1 : -1) * (row->n_gapo + row->n_gape), cigar, FQB_MAX_CIGAR, w);
        if (w.lane == 0) {
            if (nc < 0) retry[atomicAdd(n_retry, 1u)] = r;
            else {
                for (int k = 0; k < FQB_MAX_CIGAR; ++k) row->cigar[k] = cigar[k];
                row->pos = pos; row->n_cigar = (uint8_t)nc; row->has_cigar = 1;
            }
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

// as many warp slabs as the byte pool holds, at most eight blocks per SM (pool.n_blocks = 2 per SM)
static int warp_blocks(const DpPool &pool, int wpb = kWarpsPerBlock) {
    const size_t pool_bytes = (size_t)pool.n_blocks * kDpThreads * pool.bytes_per_lane;
    size_t nb = pool_bytes / (kWarpSlab * wpb);
    const size_t want = (size_t)pool.n_blocks * 16 / wpb;           // 32 warps per SM
    if (nb > want) nb = want;
    return (int)(nb < 1 ? 1 : nb);
}
// warps per block: kWarpsPerBlock while a warp's share of shared memory is small; when it is large (long reads: the trace-back
// of the refinement is 128 bytes per row) single-warp blocks pack an SM's shared memory without rounding losses
static int warps_per_block(size_t smem_per_warp) { return smem_per_warp * kWarpsPerBlock <= 56u * 1024u ? kWarpsPerBlock : 1; }

// ctr: [0] n_list [1] cursor [2] n_retry [3] retry cursor (device words)
void launch_sw(const DpView &v, const SwParams *sp, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, void *huge_mem,
               int max_read_len, cudaStream_t s) {
    sw_classify_kernel<<<(v.n_reads / 2 + 255) / 256, 256, 0, s>>>(v, sp, list, ctr);
    int ints = 2 * kSwSmemInts;                                          // H and E rows of a <= 702-column window (reverse pass) ...
    if (ints < 17 * (max_read_len + 1)) ints = 17 * (max_read_len + 1);  // ... then the local region's global alignment: row constants + trace-back (64 B per row)
    const int ref_cap = kSwSmemInts;                                     // reference codes of the window, one byte each
    const int ops_cap = 768;                                             // path ops of the local region
    const int wpb = warps_per_block((size_t)ints * 4 + ref_cap + ops_cap);
    const size_t smem = ((size_t)ints * 4 + (size_t)(ref_cap + ops_cap)) * wpb;
    // the rare very wide windows: buffers carved from huge_mem (launch_sw_huge_bytes())
    HugeBuf hb;
    char *hm = static_cast<char *>(huge_mem);
    hb.ctr = reinterpret_cast<uint32_t *>(hm); hm += 64;
    uint32_t *huge_list = reinterpret_cast<uint32_t *>(hm); hm += kHugeMaxPairs * 4;
    hb.jobs = reinterpret_cast<HugeJob *>(hm); hm += kHugeMaxJobs * sizeof(HugeJob);
    hb.slice_job = reinterpret_cast<uint32_t *>(hm); hm += (size_t)kHugeMaxSlices * 4;
    hb.slice_best = reinterpret_cast<ScanBest *>(hm);
    uint32_t *n_huge = hb.ctr + 4;
    cudaMemsetAsync(hb.ctr, 0, 64, s);
    if (getenv("FQB_DP_NO_WARP"))        // debugging aid: everything through the one-alignment-per-lane kernels
        sw_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, sp, pool, list, ctr, ctr + 1, err, huge_list, n_huge);
    else {
        cudaFuncSetAttribute(sw_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sw_warp_kernel<<<warp_blocks(pool, wpb), wpb * 32, smem, s>>>(v, sp, pool, list, ctr, ctr + 1, ints, ref_cap, ops_cap, retry, ctr + 2);
        sw_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, sp, pool, retry, ctr + 2, ctr + 3, err, huge_list, n_huge);
    }
    // windows wider than the per-lane scratch: sliced forward scan over all SMs, then the pair is finished by one warp
    sw_huge_prepare_kernel<<<1, kHugeMaxPairs, 0, s>>>(v, sp, huge_list, n_huge, hb, ref_cap, err);
    const int scan_ints = 2 * (kScanCore + kScanOverlap + 2), scan_ref = kScanCore + kScanOverlap;
    const size_t scan_smem = (size_t)scan_ints * 4 * 4 + (size_t)scan_ref * 4;
    cudaFuncSetAttribute(sw_huge_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem);
    sw_huge_scan_kernel<<<pool.n_blocks, 4 * 32, scan_smem, s>>>(v, sp, hb, scan_ints, scan_ref);
    cudaFuncSetAttribute(sw_huge_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sw_huge_finish_kernel<<<8 * kWarpsPerBlock / wpb, wpb * 32, smem, s>>>(v, sp, pool, huge_list, n_huge, hb, ints, ref_cap, ops_cap, err);
}
size_t launch_sw_huge_bytes() {
    return 64 + kHugeMaxPairs * 4 + kHugeMaxJobs * sizeof(HugeJob) + (size_t)kHugeMaxSlices * 4 + (size_t)kHugeMaxSlices * sizeof(ScanBest);
}
void launch_refine(const DpView &v, const DpPool &pool, uint32_t *list, uint32_t *retry, uint32_t *ctr, uint32_t *err, int max_read_len, cudaStream_t s) {
    refine_classify_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v, list, ctr);
    const int wcols = max_read_len + 16 + 1;                             // window columns + 1 kept in shared memory
    int ints = 6 * wcols;                                                // row-chunk form: M/I/D rows, current and previous
    const int trace_ints = (wcols <= 129 ? 17 : 33) * (max_read_len + 1); // wavefront form: row constants + trace-back, 64 or 128 B per row
    if (ints < trace_ints) ints = trace_ints;
    const int ref_cap = (wcols + 3) & ~3;                                // padded to a word
    const int ops_cap = (wcols + max_read_len + 2 + 3) & ~3;
    const int wpb = warps_per_block((size_t)ints * 4 + ref_cap + ops_cap);
    const size_t smem = ((size_t)ints * 4 + (size_t)(ref_cap + ops_cap)) * wpb;
    cudaFuncSetAttribute(refine_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (getenv("FQB_DP_NO_WARP"))
        refine_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, pool, list, ctr, ctr + 1, err);
    else {
        refine_warp_kernel<<<warp_blocks(pool, wpb), wpb * 32, smem, s>>>(v, pool, list, ctr, ctr + 1, ints, ref_cap, ops_cap, retry, ctr + 2);
        refine_kernel<<<pool.n_blocks, kDpThreads, 0, s>>>(v, pool, retry, ctr + 2, ctr + 3, err);
    }
    finish_kernel<<<(v.n_reads + 255) / 256, 256, 0, s>>>(v);
}

}  // namespace fqb
