// Deterministic synthetic reduced reference + read simulator.
// hs37d5 / dbSNP are unavailable offline, so flanks are i.i.d. random bases with
// the real marker-set shape (1000 long + 9000 short + 100 X + 97 Y by default).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "fq_index.h"

namespace fqb {

struct SynthMarker {
    std::string chrom;
    int32_t pos;          // 1-based on the synthetic genome
    char ref, alt;
    double af;
    bool is_long;
    int32_t flank;        // flank length used for this marker
    std::string flank_seq;  // 2*flank+1 bases, REF at the centre
    std::vector<uint8_t> gc;  // 100-bp-window GC counts, one per flank base (.gc record)
    int32_t extra_snp_pos;    // one extra dbSNP site inside the flank (1-based genome pos)
};

struct SynthRefConfig {
    uint64_t seed = 0x5EED0001ull;
    int n_long = 1000, n_short = 9000, n_x = 100, n_y = 97;
    int flank_short = 250, flank_long = 1000;
    int spacing = 3000;   // distance between neighbouring markers on a chromosome
    int n_dup = 0;        // segments of 200 bases copied from one marker's flank into another's (exact repeats: REPEAT-type reads)
};

struct SynthRef {
    SynthRefConfig cfg;
    std::vector<std::string> chrom_names;          // "1".."22","X","Y" in genome order
    std::vector<std::string> chrom_seq;
    std::vector<SynthMarker> markers;               // VCF (genome) order
    std::vector<int> flank_order;                   // marker indices in flank-FASTA order (chrom string order, then pos)
};

void synth_reference(const SynthRefConfig &cfg, SynthRef &out);
// Writes genome.fa(+.fai,.amb), markers.vcf, dbsnp.vcf under dir; returns false on IO error.
bool synth_write_reference_inputs(const SynthRef &ref, const std::string &dir, std::string &err);
// Writes <prefix>.FASTQuick.fa + .gc/.SelectedSite.vcf/.dbSNP.subset.vcf/.bed/.param next to the
// binary index files (what RefBuilder::PrepareRefSeq + runIndex leave behind).
bool synth_write_index_side_files(const SynthRef &ref, const std::string &genome_path,
                                  const std::string &dbsnp_path, const std::string &prefix, std::string &err);
std::vector<FlankSeq> synth_flanks(const SynthRef &ref);

struct SynthReadConfig {
    uint64_t seed = 0x5EED0002ull;
    int read_len = 100;
    double f_on = 1.0;          // fraction of pairs drawn from marker flanks
    double sub_rate = 0.01, ins_rate = 0.001, del_rate = 0.001;
    int max_indel_len = 1;
    double n_rate = 0.005;
    double isize_mean = 350, isize_sd = 40;
    double bad_tail_rate = 0.15;  // reads whose 3' qualities collapse (exercises bwa_trim_read under --q 15)
};

// Fills bases/quals (ASCII, n_pairs rows of read_len bytes per end) for pairs
// [first_pair, first_pair + n_pairs).  Thread-order independent (per-pair RNG).
void synth_reads(const SynthRef &ref, const SynthReadConfig &cfg, int64_t first_pair, int64_t n_pairs,
                 uint8_t *bases1, uint8_t *quals1, uint8_t *bases2, uint8_t *quals2, int n_threads);

}  // namespace fqb
