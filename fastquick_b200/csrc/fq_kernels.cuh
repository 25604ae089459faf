// CUDA kernels of the align hot path (sm_100a).  Per-lane logic lives in
// fq_device_core.cuh; this file holds the __global__ wrappers: work distribution,
// shared-memory staging and warp-cooperative pieces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fq_device_core.cuh"

namespace fqb {

// read r = 2 * pair + end  (the order the reference's drand48 stream visits reads,
// src/BwtMapper.cpp:744-765)
struct BatchView {
    int n_reads;            // 2 * n_pairs
    int stride_in;          // bytes per read in the ASCII input arrays (quality rows in packed form)
    int packed_stride;      // 0: bases_in holds ASCII; else bytes per read of the 2-bit rows (multiple of 16), see fqb_pack_reads
    int lpad;               // bytes per read in codes/qual
    const uint8_t *bases_in[2];
    const uint8_t *quals_in[2];
    const int32_t *lens_in[2];   // may be null: every read is stride_in long
    uint8_t *codes;         // nt4, read orientation, full_len bytes
    uint8_t *qual;          // ASCII (phred+33 after the optional -31)
    int32_t *len;           // after bwa_trim_read
    int32_t *full_len;
    uint8_t *filtered;
    uint8_t *n_ambig;       // bases > 3 among the first len (bwt_match_gap's too-many-N test)
    uint32_t *work;         // reads that go through the aligner
    uint32_t *n_work;
};

constexpr int kPrepShortFlag = 12;   // word of the batch counters (BatchView::n_work[...]) prep_kernel raises for a read shorter than the k-mer filter's 96 bases
struct PrepParams {
    int trim_qual, kmer_thresh, is_il13;
    const uint8_t *roll;    // 6 x 2^29-byte bitmaps, or null when kmer_thresh == 0
};

struct WidthView {
    uint32_t *w;            // [n_reads][2][wstride] packed widths
    uint32_t *sw;           // [n_reads][2][sstride] packed seed widths
    int wstride, sstride;
};

struct SearchParams {
    DevBwt bwt[2];
    SearchOpt opt;
    const int32_t *maxdiff;  // per read length
    int seed_len_opt;
    const uint32_t *work; const uint32_t *n_work;
    uint32_t *cursor;        // work-queue head
    uint4 *arena; uint32_t arena_cap;
    Hit *aln; int aln_cap;
    const int32_t *aln_row;  // optional: read -> row of aln (overflow pass); null = row r
    int32_t *n_aln;
    uint32_t *overflow; uint32_t *n_overflow;   // reads to redo with the big arena
    uint32_t *pops_out;                         // optional (experiments): stack pops of each read
    unsigned long long *counters;               // [0] pops, [1] rank-query pairs, [2] reference-equivalent occ-block touches (N_blk)
};

void launch_prep(const BatchView &b, const PrepParams &p, cudaStream_t s);
void launch_order(const BatchView &b, const WidthView &wv, const uint32_t *work, const uint32_t *n_work, int max_work,
                  uint32_t *bins, uint32_t *out, cudaStream_t s, const uint32_t *cost_hint = nullptr);
void launch_width(const BatchView &b, const WidthView &wv, const DevBwt bwt[2], int seed_len, const uint32_t *work,
                  const uint32_t *n_work, int max_work, unsigned long long *counters, cudaStream_t s);
// returns the number of thread blocks launched (persistent grid); heads16 selects the 16-bit head table
int search_grid_blocks(int n_buckets, bool heads16, int device);
void launch_search(const BatchView &b, const WidthView &wv, const SearchParams &p, bool heads16, bool free_list, int n_blocks, cudaStream_t s);
#ifndef FQB_SEARCH_THREADS
#define FQB_SEARCH_THREADS 128
#endif
constexpr int kSearchThreads = FQB_SEARCH_THREADS;      // 5 blocks of 128 per SM; a development build may pass another block size

}  // namespace fqb
