// C++ host surface of the align stage, mirroring the reference's classes so that `runAlign`
// (src/FASTQuick.cpp:159-491) reads the same: BwtIndexer(thresh) + LoadIndex(prefix), and BwtMapper whose
// constructor runs the whole stage (src/BwtMapper.h:38-110, src/BwtIndexer.h:53-229, src/StatCollector.h:63-228).
// Everything below forwards to the C ABI in include/fastquick_b200.h; there is no CPU implementation here.
#pragma once
#include <string>
#include <vector>

#include "../../../include/fastquick_b200.h"

namespace fqb200 {

// gap_opt_t / pe_opt_t with the reference's field names (libbwa/bwtaln.h:98-130)
struct gap_opt_t {
    int s_mm = 3, s_gapo = 11, s_gape = 4;
    int mode = 0x01 | 0x02;
    int indel_end_skip = 5, max_del_occ = 10, max_entries = 2000000;
    double fnr = 0.02, frac = 1.0;
    int max_diff = -1, max_gapo = 1, max_gape = 6;
    int max_seed_diff = 2, seed_len = 32;
    int n_threads = 4, max_top2 = 30, trim_qual = 0;
    int flank_len = 250, flank_long_len = 1000;
    unsigned num_variant_short = 9000, num_variant_long = 1000;
    char cal_dup = 1;
    std::string RG;
    char out_bam = 1;
    int read_len = 151;
};
struct pe_opt_t {
    int max_isize = 500, force_isize = 0;
    unsigned max_occ = 100000;
    int n_multi = 3, N_multi = 10, type = 1, is_sw = 1, is_preload = 0;
    double ap_prior = 1e-5;
};

void notice(const char *fmt, ...);
void warning(const char *fmt, ...);
[[noreturn]] void error(const char *fmt, ...);

class BwtIndexer {
public:
    explicit BwtIndexer(int thresh = 3) { RollParam.thresh = thresh; }
    bool LoadIndex(std::string &NewRef);          // records the prefix; the tables are uploaded when the mapper creates its engine
    struct { int kmer_size = 32, read_step_size = 1, thresh = 3; } RollParam;
    std::string RefPath, IndexPrefix;             // REFERENCE_PATH of <prefix>.param, "<prefix>"
};

class FileStatCollector {
public:
    long long NumRead = 0, NumBase = 0, HashFiltered = 0, TotalFiltered = 0, BwaUnmapped = 0, TotalMAPQ = 0, TotalRetained = 0;
    std::string FileName1, FileName2;
    FileStatCollector() = default;
    FileStatCollector(const char *f1, const char *f2) : FileName1(f1), FileName2(f2) {}
    explicit FileStatCollector(const char *f1) : FileName1(f1), FileName2(f1) {}      // src/StatCollector.h:58
};

// StatCollector: accumulation lives on the GPU inside the engine; this object is the handle-side view
class StatCollector {
public:
    explicit StatCollector(fqb_handle *h = nullptr) : h_(h) {}
    void Attach(fqb_handle *h) { h_ = h; }
    int RestoreVcfSites(const std::string &RefPath, const gap_opt_t *opt);   // fqb_stats_open
    int ProcessCore(const std::string &statPrefix, const gap_opt_t *opt);   // fqb_stats_finish
private:
    fqb_handle *h_;
};

class BwtMapper {
public:
    BwtMapper(BwtIndexer &BwtIndex, const std::string &FQList, const std::string &Fastq_1, const std::string &Fastq_2,
              const std::string &Prefix, const std::string &RefPath, const pe_opt_t *popt, gap_opt_t *opt,
              const std::string &targetRegionPath, int device = 0) : BwtMapper(BwtIndex, FQList, Fastq_1, Fastq_2, Prefix, RefPath, popt, opt, targetRegionPath, std::vector<int>{device}) {}
    // the same stage on several GPUs of this host (`--devices 0,1,...`): batches of 262,144 pairs go round-robin to the devices
    // (src/BwtMapper.h:36-37 is the unit), the index is replicated, and the library hands the drand48 position / last_ii from
    // batch to batch and merges the statistics at the end (fqb_comm_*); every output file equals the single-GPU run's
    BwtMapper(BwtIndexer &BwtIndex, const std::string &FQList, const std::string &Fastq_1, const std::string &Fastq_2,
              const std::string &Prefix, const std::string &RefPath, const pe_opt_t *popt, gap_opt_t *opt,
              const std::string &targetRegionPath, const std::vector<int> &devices);
    ~BwtMapper();
    bool PairEndMapper(const std::string &fq1, const std::string &fq2, const gap_opt_t *opt, FileStatCollector &FSC);
    bool SingleEndMapper(const std::string &fq1, const gap_opt_t *opt, FileStatCollector &FSC);
private:
    bool PairEndMapperSharded(const std::string &fq1, const std::string &fq2, const gap_opt_t *opt, FileStatCollector &FSC);
    fqb_handle *h_ = nullptr;              // rank 0: owns the BAM file and writes the statistics files
    std::vector<fqb_handle *> hs_;         // one engine per device (hs_[0] == h_)
    uint64_t pair_base_ = 0;               // global index of the next pair (sharded runs)
    StatCollector collector;
    std::string prefix_;
    bool bam_out_ = false;
};

int runAlign(int argc, char **argv);

}  // namespace fqb200
