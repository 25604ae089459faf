// Launchers of the statistics rows (fq_stats_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "fq_device_stats.cuh"

namespace fqb {

struct StatsView {
    int n_reads, lpad;
    const uint8_t *codes, *qual;   // nt4 codes / ASCII qualities (phred+33), read orientation, full_len bytes
    fqb_read_t *rows;
    PairStat *pstat;               // [n_pairs]
    const ContigDev *ctg; int n_ctg;
    uint32_t pair_base;            // global index of pair 0 of this batch (arrival order of pile-up tuples, contig first-touch)
    int cal_dup;
    const uint8_t *pac;
};

void launch_classify(const StatsView &v, const StatAccum &A, cudaStream_t s);
void launch_bases(const StatsView &v, const BaseTables &T, cudaStream_t s);

}  // namespace fqb
