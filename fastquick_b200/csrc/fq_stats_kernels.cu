// CUDA kernels of the statistics rows (a12, a13), sm_100a.
//   classify_kernel : StatCollector::AddAlignment / ProcessPairStatus per pair (thread per pair)
//   bases_kernel    : AddSingleAlignment's per-base walk, one WARP per read: lanes take consecutive read
//                     offsets, so per-site depth counters are hit with coalesced red.global.add on consecutive
//                     words; the quality / cycle histograms are privatised in shared memory and aggregated per
//                     warp with __match_any_sync before they touch an atomic
#include "fq_stats_kernels.cuh"

namespace fqb {

#define FULL_MASK 0xffffffffu

__global__ void __launch_bounds__(128) classify_kernel(StatsView v, StatAccum A) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < (uint32_t)v.n_reads / 2;
    unsigned long long c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    if (valid) {
        fqb_read_t r0 = v.rows[2 * p], r1 = v.rows[2 * p + 1];
        PairStat o;
        classify_pair(v.ctg, v.n_ctg, r0, r1, v.pair_base + p, v.cal_dup, A, o);
        if (o.demoted[0]) v.rows[2 * p].type = kTypeNoMatch;        // AddAlignment mutates p/q before SetSamRecord sees them
        if (o.demoted[1]) v.rows[2 * p + 1].type = kTypeNoMatch;
        v.pstat[p] = o;
        c3 = o.both_filtered; c4 = o.both_unmapped; c5 = o.low_mapq; c6 = o.retained; c7 = (unsigned long long)(r0.full_len + r1.full_len);
    }
    // FileStatCollector counters of the main-thread loop (src/BwtMapper.cpp:2059-2076): one atomic per warp and counter
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c3 += __shfl_xor_sync(FULL_MASK, c3, d); c4 += __shfl_xor_sync(FULL_MASK, c4, d); c5 += __shfl_xor_sync(FULL_MASK, c5, d);
        c6 += __shfl_xor_sync(FULL_MASK, c6, d); c7 += __shfl_xor_sync(FULL_MASK, c7, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c3) atomicAdd(A.scalars + 3, c3);
        if (c4) atomicAdd(A.scalars + 4, c4);
        if (c5) atomicAdd(A.scalars + 5, c5);
        if (c6) atomicAdd(A.scalars + 6, c6);
        if (c7) atomicAdd(A.scalars + 7, c7);
    }
}

// smem histogram bump with warp aggregation: lanes that share a bin elect one leader
__device__ __forceinline__ void hist_add(unsigned int *h, bool on, uint32_t bin) {
    const unsigned act = __ballot_sync(FULL_MASK, on);
    if (!on) return;
    const unsigned peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(h + bin, (unsigned)__popc(peers));
}

__global__ void __launch_bounds__(256) bases_kernel(StatsView v, BaseTables T) {
    __shared__ unsigned int s_emp[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&s_emp[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < (uint32_t)v.n_reads; r += n_warps) {
        const PairStat &ps = v.pstat[r >> 1];
        if (!ps.add[r & 1]) continue;
        const fqb_read_t s = v.rows[r];
        const uint8_t *fwd = v.codes + (size_t)r * v.lpad, *ql = v.qual + (size_t)r * v.lpad;
        const int full = s.full_len;
        // lanes walk read offsets (alignment orientation) lane, lane+32, ...
        for (int base = 0; base < full; base += 32) {
            const int off = base + lane;
            bool is_m = false;
            uint32_t x = 0;
            if (off < full) {
                if (s.has_cigar) {
                    int y = 0; uint32_t xx = s.pos;
                    for (int k = 0; k < s.n_cigar; ++k) {
                        const int op = s.cigar[k] >> 14, cl = s.cigar[k] & 0x3fff;
                        if (op == kOpM) { if (off < y + cl) { is_m = true; x = xx + (uint32_t)(off - y); break; } y += cl; xx += cl; }
                        else if (op == kOpD) xx += cl;
                        else { if (off < y + cl) break; y += cl; }          // I or S: read bases without a reference position
                    }
                } else { is_m = off < s.len; x = s.pos + (uint32_t)off; }
            }
            uint32_t rb = 4, q = 0, st = kSiteNone;
            int cycle = 0;
            if (is_m) {
                const int fo = s.strand ? full - 1 - off : off;            // offset in the read as sequenced
                rb = fwd[fo]; if (s.strand && rb < 4) rb = 3 - rb;
                q = (uint32_t)ql[fo] - 33u;
                cycle = fo;                                                 // tmpCycle: sequencing cycle of this base
                st = T.site[x];
            }
            const bool in_site = is_m && (st & kSiteMask) != kSiteNone;
            if (is_m && (st & kSiteMarker)) {                               // UpdateInfoVecAtMarker
                const uint32_t slot = atomicAdd(T.n_tuples, 1u);
                if (slot < T.tuple_cap) {
                    PileupTuple t;
                    t.marker = (uint32_t)T.marker[x];
                    t.key_hi = v.pair_base + (r >> 1); t.key_lo = ((r & 1) << 16) | (uint32_t)off;
                    t.base = (uint8_t)rb; t.qual = (uint8_t)q; t.mapq = (uint8_t)(s.mapQ + 33); t.strand = s.strand; t.cycle = cycle;
                    T.tuples[slot] = t;
                }
            }
            if (in_site) {                                                  // UpdateInfoVecAtRegularSite
                const uint32_t sid = st & kSiteMask;
                atomicAdd(T.depth + sid, 1u);
                if ((int8_t)q >= 20) { atomicAdd(T.q20 + sid, 1u); if ((int8_t)q >= 30) atomicAdd(T.q30 + sid, 1u); }
            }
            // StatVecDistUpdate: EmpRepDist[q], EmpCycleDist[cycle], mis* when the base disagrees with the reference
            const uint32_t refb = in_site ? (v.pac[x >> 2] >> ((~x & 3) << 1) & 3) : 0;
            const bool mis = in_site && rb < 4 && refb != rb && !(st & kSiteDbsnp);
            hist_add(s_emp[0], in_site, q & 255);
            hist_add(s_emp[2], in_site, (uint32_t)cycle & 255);
            hist_add(s_emp[1], mis, q & 255);
            hist_add(s_emp[3], mis, (uint32_t)cycle & 255);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
        unsigned int c = (&s_emp[0][0])[i];
        if (c) atomicAdd(T.emp + i, (unsigned long long)c);
    }
}

void launch_classify(const StatsView &v, const StatAccum &A, cudaStream_t s) {
    const int np = v.n_reads / 2;
    classify_kernel<<<(np + 127) / 128, 128, 0, s>>>(v, A);
}
void launch_bases(const StatsView &v, const BaseTables &T, cudaStream_t s) {
    int blocks = 148 * 8;
    if ((long long)blocks * 8 > v.n_reads) blocks = (v.n_reads + 7) / 8;
    if (blocks < 1) blocks = 1;
    bases_kernel<<<blocks, 256, 0, s>>>(v, T);
}

}  // namespace fqb
