// CUDA kernels of the align hot path, hand-written for sm_100a.
//   prep_kernel    a1+a2: nt4 encode, bwa_trim_read, k-mer pre-filter (warp per read)
//   width_kernel   a3:    bwt_cal_width (thread per read x {strand, full|seed})
//   search_kernel  a4+a5: bwt_match_gap, one read per lane, persistent lanes with a
//                         warp-aggregated work queue; score-bucket heads in shared
//                         memory, entry arena in (L2-backed) global memory
#include <cstdlib>
#include "fq_kernels.cuh"

#include <climits>
#include "../../include/fastquick_b200.h"

#include "fq_kmer.cuh"

namespace fqb {

#define FULL_MASK 0xffffffffu

// nst_nt4_table (libbwa/bntseq.c:38-55)
__device__ __forceinline__ uint32_t nt4_code(uint32_t ch) {
    uint32_t u = ch & 0xDFu;                       // fold case for letters
    uint32_t c = 4;
    c = (u == 'A') ? 0u : c;
    c = (u == 'C') ? 1u : c;
    c = (u == 'G') ? 2u : c;
    c = (u == 'T') ? 3u : c;
    c = (ch == '-') ? 5u : c;
    return c;
}

__global__ void __launch_bounds__(256) prep_kernel(BatchView b, PrepParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < b.n_reads; r += n_warps) {
        const int e = r & 1, pr = r >> 1;
        if (e && !b.bases_in[1]) {          // single-end input: the second read of every slot is absent
            if (lane == 0) { b.n_ambig[r] = 0; b.len[r] = 0; b.full_len[r] = 0; b.filtered[r] = 1; }
            continue;
        }
        const uint8_t *bi = (e ? b.bases_in[1] : b.bases_in[0]) + (size_t)pr * b.stride_in;
        const uint8_t *qi = (e ? b.quals_in[1] : b.quals_in[0]) + (size_t)pr * b.stride_in;
        const int32_t *li = e ? b.lens_in[1] : b.lens_in[0];
        const int full = li ? li[pr] : b.stride_in;
        uint8_t *co = b.codes + (size_t)r * b.lpad;
        uint8_t *qo = b.qual + (size_t)r * b.lpad;
        const int qadj = p.is_il13 ? 31 : 0;
        uint32_t codes[FQB_MAX_READ_LEN / 32];
        if (b.packed_stride) {
            // packed input (fqb_pack_reads): 2 bits per base, 16 bases per 32-bit word, rows aligned to 16 bytes -- the row is
            // fetched with 128-bit loads by the first lanes and handed round with shuffles; bit 7 of a quality byte marks a
            // base that is not A/C/G/T, whose 2-bit field then holds nt4 code - 4 (0 = N and friends, 1 = '-')
            const uint4 *row = reinterpret_cast<const uint4 *>((e ? b.bases_in[1] : b.bases_in[0]) + (size_t)pr * b.packed_stride);
            const int n_vec = b.packed_stride >> 4;              // <= 4 for reads up to 256 bases
            uint4 v = make_uint4(0, 0, 0, 0);
            if (lane < n_vec) v = __ldg(row + lane);
#pragma unroll
            for (int t = 0; t < FQB_MAX_READ_LEN / 32; ++t) {
                const int j = lane + 32 * t;
                // bases 32t .. 32t+31 live in words 2t and 2t+1 = components (2t & 3), (2t & 3) + 1 of vector t / 2
                const uint32_t lo = (t & 1) ? v.z : v.x, hi = (t & 1) ? v.w : v.y;
                const uint32_t w0 = __shfl_sync(FULL_MASK, lo, t >> 1), w1 = __shfl_sync(FULL_MASK, hi, t >> 1);
                codes[t] = 0;
                if (j < full) {
                    const uint32_t q = qi[j];
                    const uint32_t c2 = ((lane < 16 ? w0 : w1) >> (2 * (lane & 15))) & 3u;
                    codes[t] = (q & 0x80u) ? 4u + c2 : c2;
                    co[j] = (uint8_t)codes[t];
                    qo[j] = (uint8_t)((q & 0x7fu) - qadj);
                }
            }
        } else {
#pragma unroll
        for (int t = 0; t < FQB_MAX_READ_LEN / 32; ++t) {
            int j = lane + 32 * t;
            codes[t] = 0;
            if (j < full) {
                codes[t] = nt4_code(bi[j]);
                co[j] = (uint8_t)codes[t];
                qo[j] = (uint8_t)(qi[j] - qadj);
            }
        }
        }
        // ---- k-mer pre-filter: IsReadInHashByCountMoreChunck (src/BwtIndexer.cpp:441-456).
        // kmer = (kmer << 2) | code over 32 bases == OR of shifted codes (N = 4 bleeds upward).
        bool filtered = false;
        if (p.kmer_thresh != 0) {
            // The reference's filter always reads bases 0..95 of its per-slot buffer (src/BwtIndexer.cpp:443-450): for a
            // shorter read it sees what an earlier read left there.  There is nothing to match: flag the batch.
            if (full < 96 && lane == 0) b.n_work[kPrepShortFlag] = 1;
            uint64_t kmer[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                uint64_t v = (uint64_t)codes[t] << (2 * (31 - lane));
                uint32_t lo = __reduce_or_sync(FULL_MASK, (uint32_t)v);
                uint32_t hi = __reduce_or_sync(FULL_MASK, (uint32_t)(v >> 32));
                kmer[t] = ((uint64_t)hi << 32) | lo;
            }
            uint32_t bit = 0;
            if (lane < 18) {
                int ch = lane / 6, tb = lane - 6 * ch;
                uint64_t km = ch == 0 ? kmer[0] : ch == 1 ? kmer[1] : kmer[2];
                uint32_t x = shrink_kmer(km, tb);
                bit = (p.roll[((size_t)tb << 29) + (x >> 3)] >> (x & 7)) & 1u;
            }
            int hits = __popc(__ballot_sync(FULL_MASK, bit));
            filtered = hits < p.kmer_thresh;
        }
        // ---- bwa_trim_read (libbwa/bwaseqio.c:75-88), suffix sums by warp scan
        int len = full;
        if (p.trim_qual >= 1) {
            int carry = 0, best = 0, best_l = full - 1;
            for (int base = full - 1; base >= 34; base -= 32) {
                int l = base - lane;
                bool valid = l >= 34;
                int v = valid ? p.trim_qual - ((int)(qi[l] & (b.packed_stride ? 0x7fu : 0xffu)) - qadj - 33) : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(FULL_MASK, v, d);
                    if (lane >= d) v += t;
                }
                int s = v + carry;
                unsigned neg = __ballot_sync(FULL_MASK, valid && s < 0);
                int first_neg = neg ? __ffs(neg) - 1 : 32;
                bool ok = valid && lane < first_neg;
                long long key = ok ? (((long long)s << 8) | (long long)(63 - lane)) : LLONG_MIN;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    long long o = __shfl_xor_sync(FULL_MASK, key, d);
                    key = o > key ? o : key;
                }
                if (key != LLONG_MIN) {
                    int cs = (int)(key >> 8), cl = 63 - (int)(key & 0xff);
                    if (cs > best) { best = cs; best_l = base - cl; }
                }
                carry = __shfl_sync(FULL_MASK, s, 31);
                if (neg) break;
            }
            len = best_l + 1;
        }
        int n_ambig = 0;
#pragma unroll
        for (int t = 0; t < FQB_MAX_READ_LEN / 32; ++t)
            n_ambig += __popc(__ballot_sync(FULL_MASK, codes[t] > 3 && lane + 32 * t < len));
        if (lane == 0) {
            b.n_ambig[r] = (uint8_t)(n_ambig > 255 ? 255 : n_ambig);
            b.len[r] = len;
            b.full_len[r] = full;
            b.filtered[r] = filtered ? 1 : 0;
            if (!filtered) b.work[atomicAdd(b.n_work, 1u)] = (uint32_t)r;
        }
    }
}

void launch_prep(const BatchView &b, const PrepParams &p, cudaStream_t s) {
    int blocks = (b.n_reads + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    prep_kernel<<<blocks, 256, 0, s>>>(b, p);
}

// ---------------------------------------------------------------------------
struct WidthParams { DevBwt bwt[2]; };

__global__ void __launch_bounds__(128) width_kernel(BatchView b, WidthView wv, WidthParams wp, int seed_len,
                                                     const uint32_t *work, const uint32_t *n_work, unsigned long long *counters) {
    // a warp handles 32 reads for one part, so lanes run loops of the same trip count
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t wi = (idx >> 7) * 32 + (idx & 31);
    const int part = (idx >> 5) & 3;
    if (wi >= *n_work) return;
    const uint32_t r = work[wi];
    const int len = b.len[r];
    const int a = part & 1;
    const uint8_t *fwd = b.codes + (size_t)r * b.lpad;
    const bool seed = part >= 2;
    if (seed && len <= seed_len) return;
    const int first = seed ? len - seed_len : 0, n = seed ? seed_len : len;
    uint32_t *out = seed ? wv.sw + ((size_t)r * 2 + a) * wv.sstride : wv.w + ((size_t)r * 2 + a) * wv.wstride;
    uint32_t touches = a == 0 ? cal_width(wp.bwt[0], fwd, len, 0, first, n, out) : cal_width(wp.bwt[1], fwd, len, 1, first, n, out);
    if (counters) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) touches += __shfl_xor_sync(__activemask(), touches, d);
        if ((threadIdx.x & 31) == 0) atomicAdd(counters + 2, (unsigned long long)touches);
    }
}

void launch_width(const BatchView &b, const WidthView &wv, const DevBwt bwt[2], int seed_len, const uint32_t *work,
                  const uint32_t *n_work, int max_work, unsigned long long *counters, cudaStream_t s) {
    WidthParams wp;
    wp.bwt[0] = bwt[0]; wp.bwt[1] = bwt[1];
    long long threads = ((long long)(max_work + 31) / 32) * 128;
    int blocks = (int)((threads + 127) / 128);
    if (blocks < 1) blocks = 1;
    width_kernel<<<blocks, 128, 0, s>>>(b, wv, wp, seed_len, work, n_work, counters);
}

// ---------------------------------------------------------------------------
// Longest-first ordering of the search queue.  The cost of bwt_match_gap grows steeply with the number of
// differences a read needs; bwt_cal_width's total lower bound (bid of the last position, the smaller of the
// two strands) is known before the search starts, so the queue is bucket-sorted by it, descending.  Reads are
// independent, so the order changes nothing but the load balance of the persistent lanes.
constexpr int kOrderBins = 16;
__device__ __forceinline__ int order_key(const BatchView &b, const WidthView &wv, uint32_t r, const uint32_t *hint) {
    if (hint) { const uint32_t k = hint[r] >> 8; return k < (uint32_t)kOrderBins ? (int)k : kOrderBins - 1; }
    const int len = b.len[r];
    if (len < 1) return 0;
    const int b0 = width_bid(wv.w[((size_t)r * 2) * wv.wstride + len - 1]), b1 = width_bid(wv.w[((size_t)r * 2 + 1) * wv.wstride + len - 1]);
    const int k = b0 < b1 ? b0 : b1;
    return k < kOrderBins ? k : kOrderBins - 1;
}
__global__ void __launch_bounds__(256) order_hist_kernel(BatchView b, WidthView wv, const uint32_t *work, const uint32_t *n_work, uint32_t *bins, const uint32_t *hint) {
    __shared__ uint32_t sh[kOrderBins];
    if (threadIdx.x < kOrderBins) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < *n_work) atomicAdd(&sh[order_key(b, wv, work[idx], hint)], 1u);
    __syncthreads();
    if (threadIdx.x < kOrderBins && sh[threadIdx.x]) atomicAdd(&bins[threadIdx.x], sh[threadIdx.x]);
}
__global__ void __launch_bounds__(256) order_scatter_kernel(BatchView b, WidthView wv, const uint32_t *work, const uint32_t *n_work,
                                                             const uint32_t *bins, uint32_t *fill, uint32_t *out, const uint32_t *hint) {
    __shared__ uint32_t sh[kOrderBins], base[kOrderBins];
    if (threadIdx.x < kOrderBins) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = idx < *n_work;
    uint32_t r = 0, rank = 0;
    int key = 0;
    if (in) { r = work[idx]; key = order_key(b, wv, r, hint); rank = atomicAdd(&sh[key], 1u); }
    __syncthreads();
    if (threadIdx.x < kOrderBins) {
        uint32_t start = 0;
        for (int k = kOrderBins - 1; k > (int)threadIdx.x; --k) start += bins[k];       // heavier bins first
        base[threadIdx.x] = start + (sh[threadIdx.x] ? atomicAdd(&fill[threadIdx.x], sh[threadIdx.x]) : 0u);
    }
    __syncthreads();
    if (in) out[base[key] + rank] = r;
}
void launch_order(const BatchView &b, const WidthView &wv, const uint32_t *work, const uint32_t *n_work, int max_work,
                  uint32_t *bins /* 2 x 16 words, zeroed here */, uint32_t *out, cudaStream_t s, const uint32_t *cost_hint) {
    cudaMemsetAsync(bins, 0, 2 * kOrderBins * 4, s);
    const int blocks = (max_work + 255) / 256 < 1 ? 1 : (max_work + 255) / 256;
    order_hist_kernel<<<blocks, 256, 0, s>>>(b, wv, work, n_work, bins, cost_hint);
    order_scatter_kernel<<<blocks, 256, 0, s>>>(b, wv, work, n_work, bins, bins + kOrderBins, out, cost_hint);
}

// ---------------------------------------------------------------------------
// gap_shadow (libbwa/bwtgap.c:81-91) for every lane of the warp that just recorded a hit, 32 width
// entries at a time; the running "++j" of the reference becomes a ballot prefix count.
template <typename Lane>
__device__ __forceinline__ void warp_shadow(Lane &lane, bool has_hit, int lane_id) {
    unsigned hm = __ballot_sync(FULL_MASK, has_hit);
    while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        unsigned long long wp_ = __shfl_sync(FULL_MASK, (unsigned long long)(uintptr_t)lane.wa(), src);
        uint32_t *wp = reinterpret_cast<uint32_t *>((uintptr_t)wp_);
        const uint32_t x = __shfl_sync(FULL_MASK, lane.hit_x, src);
        const int a = __shfl_sync(FULL_MASK, lane.a, src);
        const int ldp = __shfl_sync(FULL_MASK, lane.ldp, src);
        const uint32_t maxv = (a ? lane.bwt[0] : lane.bwt[1]).seq_len;
        uint32_t j = 0;
        for (int base = 0; base < ldp; base += 32) {
            const int p = base + lane_id;
            const bool in = p < ldp;
            const uint32_t v = in ? wp[p] : 0u, ww = width_w(v);
            const bool eq = in && ww == x;
            const unsigned em = __ballot_sync(FULL_MASK, eq);
            if (in) {
                if (ww > x) wp[p] = v - x;
                else if (eq) wp[p] = pack_width(maxv - (j + (uint32_t)__popc(em & ((1u << lane_id) - 1u)) + 1u), 1);
            }
            j += (uint32_t)__popc(em);
        }
    }
    __syncwarp();
}

// Shared memory of a block: -- kStage -- one 32-byte staging slot per lane (SearchLane::stage_slot), then the bucket heads
// (n_buckets x kSearchThreads x HeadT, one column per lane).
constexpr int kStageBytes = 32;
template <typename HeadT, bool kFreeList, int kMinBlocks = 640 / kSearchThreads, int kVar = 0>
__global__ void __launch_bounds__(kSearchThreads, kMinBlocks) search_kernel(const __grid_constant__ BatchView b, const __grid_constant__ WidthView wv,
                                                                            const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kStage = (kVar & 1) && !kFreeList;
    HeadT *heads = reinterpret_cast<HeadT *>(smem_raw + (kStage ? kSearchThreads * kStageBytes : 0));
    // search options and index descriptors are read straight from the kernel-parameter constant bank
    const DevBwt *s_bwt = p.bwt;
    const SearchOpt &s_opt = p.opt;

    const int lane_id = threadIdx.x & 31;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    SearchLane<HeadT, kFreeList, kVar> lane;
    lane.bwt = s_bwt; lane.opt = &s_opt;
    lane.arena = p.arena + gtid * p.arena_cap; lane.arena_cap = p.arena_cap;
    lane.heads = heads + threadIdx.x; lane.head_stride = blockDim.x;
    lane.out_cap = p.aln_cap;
    lane.w[0] = lane.w[1] = nullptr; lane.a = 0; lane.ldp = 0; lane.hit_x = 0;

    const uint32_t n_work = *p.n_work;
    bool active = false, exhausted = false;
    uint32_t r = 0;
    unsigned long long pops = 0, occs = 0, blks = 0;
#ifdef FQB_KSTATS
    unsigned long long ks_iter = 0, ks_step = 0, ks_exh = 0; unsigned ks_hist[9] = {0,0,0,0,0,0,0,0,0};
    // timeline of the launch (development build only): slot 16 = start, 17 = when the queue ran dry, 18 = end (globaltimer ns);
    // 32.. = warps leaving, 160.. = warp trips, 288.. = live lanes summed over those trips, all in 0.25-ms bins since the start
    const bool ks_tl = p.counters && n_work > 4096;
    auto ks_now = []() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; };
    if (ks_tl && blockIdx.x == 0 && threadIdx.x == 0) p.counters[16] = ks_now();
    bool ks_dry_seen = false;
#endif

    for (;;) {
        LaneStatus st = kLaneRunning;
        unsigned need = __ballot_sync(FULL_MASK, !active && !exhausted);
        if (need) {
            uint32_t base = 0;
            int leader = __ffs(need) - 1;
            if (lane_id == leader) base = atomicAdd(p.cursor, (uint32_t)__popc(need));
            base = __shfl_sync(FULL_MASK, base, leader);
            if (!active && !exhausted) {
                uint32_t idx = base + __popc(need & ((1u << lane_id) - 1u));
                if (idx >= n_work) exhausted = true;
                else {
                    r = p.work[idx];
                    const int len = b.len[r];
                    lane.fwd = b.codes + (size_t)r * b.lpad;
                    lane.w[0] = wv.w + ((size_t)r * 2) * wv.wstride;
                    lane.w[1] = lane.w[0] + wv.wstride;
                    bool seeded = len > p.seed_len_opt;
                    lane.sw[0] = seeded ? wv.sw + ((size_t)r * 2) * wv.sstride : nullptr;
                    lane.sw[1] = seeded ? lane.sw[0] + wv.sstride : nullptr;
                    lane.out = p.aln + (size_t)(p.aln_row ? (uint32_t)p.aln_row[r] : r) * p.aln_cap;
                    st = lane.begin(len, p.maxdiff[len], b.n_ambig[r]);
                    active = true;
                }
            }
        }
        if (__all_sync(FULL_MASK, exhausted && !active)) break;
#ifdef FQB_KSTATS
        { unsigned m1 = __ballot_sync(FULL_MASK, active && st == kLaneRunning), m2 = __ballot_sync(FULL_MASK, exhausted && !active);
          if (lane_id == 0) { ks_iter++; ks_step += __popc(m1); ks_exh += __popc(m2); ks_hist[__popc(m1) >> 2]++; }
          if (ks_tl && lane_id == 0) {
              const unsigned long long t = ks_now(), t0 = *(volatile unsigned long long *)(p.counters + 16);
              unsigned bin = t0 && t > t0 ? (unsigned)((t - t0) / 250000ull) : 0; if (bin > 127) bin = 127;
              atomicAdd(p.counters + 160 + bin, 1ull); atomicAdd(p.counters + 288 + bin, (unsigned long long)__popc(m1));
              if (m2 && !ks_dry_seen) { ks_dry_seen = true; atomicMax(p.counters + 17, t); }
          } }
#endif
        if (active && st == kLaneRunning) st = lane.step();
        warp_shadow(lane, active && st == kLaneHit, lane_id);
        if (active && (st == kLaneDone || st == kLaneOverflow)) {
            active = false;
            pops += lane.n_pops; occs += lane.n_occ; blks += lane.n_blk;
            p.n_aln[r] = st == kLaneOverflow ? -1 : lane.n_aln;
            if (p.pops_out) p.pops_out[r] = lane.n_pops;
            if (st == kLaneOverflow) p.overflow[atomicAdd(p.n_overflow, 1u)] = r;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        pops += __shfl_xor_sync(FULL_MASK, pops, d);
        occs += __shfl_xor_sync(FULL_MASK, occs, d);
        blks += __shfl_xor_sync(FULL_MASK, blks, d);
    }
    if (lane_id == 0 && p.counters) { atomicAdd(p.counters, pops); atomicAdd(p.counters + 1, occs); atomicAdd(p.counters + 2, blks); }
#ifdef FQB_KSTATS
    if (ks_tl && lane_id == 0) {
        const unsigned long long t = ks_now(), t0 = *(volatile unsigned long long *)(p.counters + 16);
        unsigned bin = t0 && t > t0 ? (unsigned)((t - t0) / 250000ull) : 0; if (bin > 127) bin = 127;
        atomicAdd(p.counters + 32 + bin, 1ull); atomicMax(p.counters + 18, t);
    }
    if (lane_id == 0 && p.counters) { atomicAdd(p.counters + 4, ks_iter); atomicAdd(p.counters + 5, ks_step); atomicAdd(p.counters + 6, ks_exh);
        for (int q = 0; q < 9; ++q) atomicAdd(p.counters + 7 + q, (unsigned long long)ks_hist[q]); }
#endif
}

static int search_occ_variant() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("FQB_SEARCH_OCC"); v = e ? atoi(e) : 1; }
    return v;
}
// Memory-path form of the fast pass (SearchLane kVar, a bit mask): 5 = pop staging through shared memory + streaming stack
// stores is the build's default (profiles/r02_search_memory_variants.md: filter-bound input +10 %, 150-base reads +3 %, the
// 100-base contract workload unchanged); FQB_SEARCH_VAR = 0 / 1 / 4 / 5 overrides it.  Read at every launch, so that
// tools/stage_ab.py can time the forms in one process on one batch.
constexpr int kSearchVarDefault = 5;
static int search_mem_variant() {
    const char *e = getenv("FQB_SEARCH_VAR");
    const int v = e ? atoi(e) : kSearchVarDefault;
    return v == 0 || v == 1 || v == 4 || v == 5 ? v : kSearchVarDefault;
}
template <int kVar> struct FastSearch {
    static constexpr int kMinBlocks = 640 / kSearchThreads;
    static size_t smem(int n_buckets) { return (size_t)n_buckets * kSearchThreads * 2 + ((kVar & 1) ? kSearchThreads * kStageBytes : 0); }
    static int per_sm(int n_buckets) {
        int v = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, search_kernel<uint16_t, false, kMinBlocks, kVar>, kSearchThreads, smem(n_buckets));
        return v;
    }
    static void launch(const BatchView &b, const WidthView &wv, const SearchParams &p, int n_blocks, cudaStream_t s) {
        const size_t sm = smem(p.opt.n_buckets);
        cudaFuncSetAttribute(search_kernel<uint16_t, false, kMinBlocks, kVar>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        search_kernel<uint16_t, false, kMinBlocks, kVar><<<n_blocks, kSearchThreads, sm, s>>>(b, wv, p);
    }
};
static int fast_search_per_sm(int var, int n_buckets) {
    return var == 5 ? FastSearch<5>::per_sm(n_buckets) : var == 4 ? FastSearch<4>::per_sm(n_buckets) : FastSearch<1>::per_sm(n_buckets);
}
static void fast_search_launch(int var, const BatchView &b, const WidthView &wv, const SearchParams &p, int n_blocks, cudaStream_t s) {
    if (var == 5) FastSearch<5>::launch(b, wv, p, n_blocks, s);
    else if (var == 4) FastSearch<4>::launch(b, wv, p, n_blocks, s);
    else FastSearch<1>::launch(b, wv, p, n_blocks, s);
}

int search_grid_blocks(int n_buckets, bool heads16, int device) {
    int per_sm = 0, n_sm = 148;
    size_t smem = (size_t)n_buckets * kSearchThreads * (heads16 ? 2 : 4);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    if (heads16 && search_mem_variant()) per_sm = fast_search_per_sm(search_mem_variant(), n_buckets);
    else if (heads16 && search_occ_variant() == 6) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<uint16_t, false, 6>, kSearchThreads, smem);
    else if (heads16 && search_occ_variant() == 8) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<uint16_t, false, 8>, kSearchThreads, smem);
    else if (heads16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<uint16_t, false>, kSearchThreads, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<uint32_t, true>, kSearchThreads, smem);
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("FQB_SEARCH_BLOCKS_PER_SM")) { int c = atoi(e); if (c >= 1 && c < per_sm) per_sm = c; }
    return n_sm * per_sm;
}

template <typename HeadT, bool kFreeList, int kMinBlocks = 640 / kSearchThreads>
static void launch_search_t(const BatchView &b, const WidthView &wv, const SearchParams &p, int n_blocks, cudaStream_t s) {
    size_t smem = (size_t)p.opt.n_buckets * kSearchThreads * sizeof(HeadT);
    cudaFuncSetAttribute(search_kernel<HeadT, kFreeList, kMinBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    search_kernel<HeadT, kFreeList, kMinBlocks><<<n_blocks, kSearchThreads, smem, s>>>(b, wv, p);
}

void launch_search(const BatchView &b, const WidthView &wv, const SearchParams &p, bool heads16, bool free_list, int n_blocks, cudaStream_t s) {
    if (heads16 && !free_list && search_mem_variant()) fast_search_launch(search_mem_variant(), b, wv, p, n_blocks, s);
    else if (heads16 && !free_list && search_occ_variant() == 6) launch_search_t<uint16_t, false, 6>(b, wv, p, n_blocks, s);
    else if (heads16 && !free_list && search_occ_variant() == 8) launch_search_t<uint16_t, false, 8>(b, wv, p, n_blocks, s);
    else if (heads16 && !free_list) launch_search_t<uint16_t, false>(b, wv, p, n_blocks, s);
    else if (heads16) launch_search_t<uint16_t, true>(b, wv, p, n_blocks, s);
    else launch_search_t<uint32_t, true>(b, wv, p, n_blocks, s);
}

}  // namespace fqb
