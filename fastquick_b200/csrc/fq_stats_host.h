// Host side of the statistics rows: the per-pac-position side tables the kernels read
// (StatCollector::RestoreVcfSites, src/StatCollector.cpp:1742-1842) and the finalisation that
// writes FASTQuick's summary files (StatCollector::ProcessCore, 2012-2028 and the Get*/SummaryOutput
// writers, 1858-2483; InsertSizeEstimator, src/InsertSizeEstimator.cpp:43-173).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/fastquick_b200.h"
#include "fq_device_stats.cuh"
#include "fq_index.h"

namespace fqb {

struct MarkerRec {
    std::string chrom_raw, chrom;      // as written / upper-cased with "CHR" stripped
    int pos = 0;
    std::string id, ref, alt, qual, filter, af;
    bool has_af = false;
};

struct StatsTables {
    std::vector<ContigDev> contigs;
    std::vector<std::string> contig_names;
    std::vector<uint32_t> site;          // [l_pac] site id | kSiteDbsnp | kSiteMarker, or kSiteNone
    std::vector<int32_t> marker_at;      // [l_pac] marker index (VCF order) or -1
    std::vector<uint8_t> site_gc;        // [n_sites]
    uint32_t n_sites = 0;
    std::vector<MarkerRec> markers;      // VCF order
    std::vector<int> marker_out_order;   // VcfTable iteration order (chrom string order, then position)
    uint64_t n_short = 0, n_long = 0, n_xy = 0;
    int chopped_read_len = 0;
    uint64_t ref_genome_size = 0, ref_N_size = 0;   // BwtIndexer::LoadContigSize (src/BwtIndexer.cpp:764-802)
    bool has_target = false; uint64_t flank_region_size = 0, target_region_size = 0;   // flankRegion.Size() after InnerJoin, targetRegion.Size()
    std::vector<std::pair<std::string, int>> genome_contigs;   // BwtIndexer::contigSize: the lines of <reference>.fai (BAM header @SQ)
};

// target_bed: --targetRegion (StatCollector::SetTargetRegion), or empty
bool build_stats_tables(const HostIndex &idx, const std::string &index_prefix, const fqb_gap_opt_t &g, const std::string &target_bed, StatsTables &out,
                        std::string &err);

// FileStatCollector (src/StatCollector.h:46-62)
struct FileCounters {
    long long NumRead = 0, NumBase = 0, TotalFiltered = 0, BwaUnmapped = 0, TotalMAPQ = 0, TotalRetained = 0;
    std::string FileName1, FileName2;
};

struct PileupColumn { std::string seq, qual; std::vector<int> cycle; std::vector<unsigned char> maq; std::vector<bool> strand; };

// Everything ProcessCore needs, gathered from the device at the end of the run.
struct StatsTotals {
    std::vector<uint32_t> depth, q20, q30;                 // [n_sites]
    std::vector<unsigned long long> emp;                   // [4][256]
    std::vector<unsigned long long> isize_dist;            // [4096]
    unsigned long long num_pcr_dup = 0, num_pair_reads = 0;
    std::vector<uint32_t> contig_ctr, contig_first;        // [n_contigs][4], [n_contigs]
    std::vector<PileupColumn> pileup;                      // [n_markers], arrival order
    std::vector<FileCounters> files;
};

// One InsertSizeTable line (or nothing) for a pair, exactly as ProcessPairStatus prints it.
void format_isize_line(const StatsTables &T, const PairStat &ps, const fqb_read_t &p, const fqb_read_t &q, const char *name, std::string &out);
void append_isize_line(const StatsTables &T, const PairStat &ps, const fqb_read_t &p, const fqb_read_t &q, const char *name, std::string &out);

// InsertSizeEstimator (src/InsertSizeEstimator.cpp:43-173) over a finished InsertSizeTable -> the AdjustedInsertSizeDist file
bool write_adjusted_isize(const std::string &table, const std::string &out_path);

// ProcessCore: writes <prefix>.{DepthDist,GCDist,EmpRepDist,EmpCycleDist,AdjustedInsertSizeDist,RawInsertSizeDist,
// SexChromInfo,Pileup,FASTQ.csv,Sequence.csv,Summary,vcf}; <prefix>.InsertSizeTable must already be complete.
bool write_summary_files(const StatsTables &T, StatsTotals &S, const fqb_gap_opt_t &g, const std::string &prefix, std::string &err);

}  // namespace fqb
