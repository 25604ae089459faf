// K-mer pre-filter tables: the six spaced 16-of-32-base masks and the device-side
// construction of the 6 x 2^32-bit membership tables that the reference keeps in <prefix>.rollhash.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fqb {

// KmerShrinkage cases 0..5 (src/BwtIndexer.h:262-315)
__device__ __forceinline__ uint32_t shrink_kmer(uint64_t kmer, int which) {
    uint32_t hi = (uint32_t)(kmer >> 32), lo = (uint32_t)kmer;
    switch (which) {
    case 0: return hi;
    case 1: return lo;
    case 2: return (hi & 0xffff0000u) | (lo & 0xffffu);
    case 3: return (uint32_t)(kmer >> 16);
    case 4: return (hi & 0xffff0000u) | (lo >> 16);
    default: return (hi << 16) | (lo & 0xffffu);
    }
}

// What BwtIndexer::Fa2Pac feeds to AddSeq2Hash (src/BwtIndexer.cpp:870-885): every flank and its reverse complement,
// with the two allele characters that follow '@' in the flank name.
struct KmerBuildView {
    const uint8_t *codes;      // nst_nt4_table codes of the concatenated flank text (>= 4 at the .amb holes)
    const int64_t *offset;     // [n_flanks + 1] start of each flank in codes
    const uint8_t *alleles;    // [n_flanks][2] nt4 codes of the allele characters; bit 7 of the first = flank handled by the special list
    int n_flanks;
    int64_t n_bases;
    uint32_t *tables;          // 6 x 2^27 words, zeroed by the caller
};
void launch_kmer_build(const KmerBuildView &v, cudaStream_t s);

// Flanks with ambiguous bases: the reference substitutes a fresh rand() % 4 at every visit (NST_NT4_TABLE, src/BwtIndexer.cpp:
// 59-61), so each (strand, table) pass walks its own string.  The host reproduces the draws (fq_index.cpp kmer_build_inputs);
// a job is one such pass: `first` = the string of the first allele's pass, `last` = the second's (they differ in the 32
// positions from the centre on); windows ending before the centre read `first`, those covering it both, later ones `last`.
struct KmerSpecialView {
    const uint8_t *codes;
    const int64_t *first, *last;
    const int32_t *len, *table;
    int n_jobs;
    uint32_t *tables;
};
void launch_kmer_build_special(const KmerSpecialView &v, cudaStream_t s);

}  // namespace fqb
