// Builds the k-mer pre-filter tables on the device instead of reading the reference's 3 GiB <prefix>.rollhash
// (BwtIndexer::AddSeq2HashCore, src/BwtIndexer.cpp:611-713, called from Fa2Pac :870-885 for each flank and its
// reverse complement; ReadRollHashTable :569-579 is what this replaces at load time).
//
// The reference rolls datum = (datum << 2) | code along the string, so the value at end position i is the OR of the
// last 32 codes shifted into place (a code >= 4 spills into its neighbour's bits; older history is shifted out).  Every
// window is therefore independent: one thread per window start, both strands.  The string it walks is the flank with
// its centre base (index len/2) replaced by each of the two alleles in turn for the windows that cover the centre; the
// reverse-complement string gets the SAME allele characters at ITS index len/2, uncomplemented.
#include "fq_kmer.cuh"

namespace fqb {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void set_bits(uint32_t *tables, uint64_t kmer) {
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const uint32_t x = shrink_kmer(kmer, t);           // byte x >> 3, bit x & 7 of table t == word x >> 5, bit x & 31
        atomicOr(tables + ((size_t)t << 27) + (x >> 5), 1u << (x & 31));
    }
}

__global__ void __launch_bounds__(kThreads) kmer_build_kernel(KmerBuildView v) {
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < v.n_bases; g += (int64_t)gridDim.x * kThreads) {
        int lo = 0, hi = v.n_flanks;                       // flank of base g: last offset <= g
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (v.offset[mid] <= g) lo = mid; else hi = mid; }
        if (v.alleles[2 * lo] & 0x80) continue;            // ambiguous bases: kmer_special_kernel
        const int64_t off = v.offset[lo];
        const int len = (int)(v.offset[lo + 1] - off), p = (int)(g - off);
        if (p + 32 > len) continue;
        const int half = len / 2;
        const int kf = half - p;                           // window slot of the forward string's centre
        const int kr = (len - 1 - half) - p;               // window slot (forward coordinates) of the reverse-complement string's centre
        uint8_t c[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) c[k] = v.codes[g + k];
        const bool fvar = kf >= 0 && kf < 32, rvar = kr >= 0 && kr < 32;
        for (int a = 0; a < 2; ++a) {
            const uint32_t al = v.alleles[2 * lo + a];
            if (a == 0 || fvar) {
                uint64_t d = 0;
#pragma unroll
                for (int k = 0; k < 32; ++k) d |= (uint64_t)((fvar && k == kf) ? al : c[k]) << (2 * (31 - k));
                set_bits(v.tables, d);
            }
            if (a == 0 || rvar) {
                uint64_t d = 0;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const uint32_t cc = c[k] < 4 ? 3u - c[k] : 4u;     // ReverseComplement maps anything else to '\0', code 4
                    d |= (uint64_t)((rvar && k == kr) ? al : cc) << (2 * k);
                }
                set_bits(v.tables, d);
            }
        }
    }
}

// one block per (flank, strand, table) pass of a flank with substituted bases; one thread per window end
__global__ void __launch_bounds__(kThreads) kmer_special_kernel(KmerSpecialView v) {
    const int job = blockIdx.x;
    const int len = v.len[job], half = len / 2, t = v.table[job];
    const uint8_t *c0 = v.codes + v.first[job], *c1 = v.codes + v.last[job];
    for (int i = 31 + threadIdx.x; i < len; i += kThreads) {
        for (int a = 0; a < 2; ++a) {
            if (a == 0 ? i >= half + 32 : i < half) continue;      // before the centre: first == last; past it the rolling value came from `last`
            const uint8_t *c = a ? c1 : c0;
            uint64_t d = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) d |= (uint64_t)c[i - 31 + k] << (2 * (31 - k));
            const uint32_t x = shrink_kmer(d, t);
            atomicOr(v.tables + ((size_t)t << 27) + (x >> 5), 1u << (x & 31));
        }
    }
}

}  // namespace

void launch_kmer_build_special(const KmerSpecialView &v, cudaStream_t s) {
    if (v.n_jobs > 0) kmer_special_kernel<<<v.n_jobs, kThreads, 0, s>>>(v);
}

void launch_kmer_build(const KmerBuildView &v, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    kmer_build_kernel<<<sms * 8, kThreads, 0, s>>>(v);
}

}  // namespace fqb
