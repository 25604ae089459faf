// Host-side model of the reduced marker-flank index that FASTQuick's align
// stage loads (BwtIndexer::LoadIndex, src/BwtIndexer.cpp:803-837 of the
// reference).  File formats: libbwa/bwtio.c:7-60 (.bwt/.sa), libbwa/bntseq.c
// (.ann/.amb), src/BwtIndexer.cpp:839-975 (.pac), :569-592 (.rollhash).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fqb {

constexpr int      kOccInterval   = 128;              // OCC_INTERVAL, libbwa/bwt.h:34
constexpr uint64_t kRollTableBytes = 1ull << 29;      // 4^16 bits / 8, src/BwtIndexer.cpp:561
constexpr int      kNumRollTables = 6;

// bwt_t (libbwa/bwt.h:42-54), reference word layout kept as loaded from disk:
// per 128 bases 4 cumulative counts followed by 8 words of 16 packed bases.
struct HostBwt {
    uint32_t primary = 0;
    uint32_t L2[5] = {0, 0, 0, 0, 0};
    uint32_t seq_len = 0;
    std::vector<uint32_t> bwt;       // interleaved occ + bases
    uint32_t sa_intv = 32;
    std::vector<uint32_t> sa;        // sa[0] == 0xffffffff
};

struct Contig {                      // bntann1_t
    int64_t offset = 0;
    int32_t len = 0;
    int32_t n_ambs = 0;
    std::string name, anno;
};
struct Hole { int64_t offset; int32_t len; char amb; };   // bntamb1_t

struct HostIndex {
    HostBwt bwt[2];                  // [0] = .bwt/.sa (forward text), [1] = .rbwt/.rsa (reversed text)
    std::vector<uint8_t> pac;        // 2 bits/base, MSB first
    int64_t l_pac = 0;
    uint32_t seed = 11;
    std::vector<Contig> contigs;
    std::vector<Hole> holes;
    std::string rollhash_path;       // 6 x 512 MiB on disk (streamed to the device), or
    std::vector<uint8_t> rollhash;   // in-memory tables when built here (6 x 2^29 bytes)
};

// Loaders return false and set err on failure.
bool load_bwt(const std::string &bwt_path, const std::string &sa_path, HostBwt &out, std::string &err);
bool load_index(const std::string &prefix, bool with_rollhash_in_memory, HostIndex &out, std::string &err);

// Inputs of the device-side k-mer table build (fq_kmer_build.cu): the flank text as nt4 codes (.pac with the .amb holes put
// back), flank boundaries, and the two allele characters after '@' in each flank name (src/BwtIndexer.cpp:873-875).
// A flank holding an ambiguous base (or an allele character outside ACGT) is flagged 0x80 in alleles[2 f] and listed in
// `special` instead: there the reference substitutes rand() % 4 at every visit (NST_NT4_TABLE, src/BwtIndexer.cpp:59-61), so
// each of its 2 strands x 6 tables walks its own string; the draws are reproduced here on the host, in the reference's order.
// false = a flank name has no "@x/y" part or a flank is shorter than 65 bases (the reference's loops assume neither).
struct KmerSpecialJob { int64_t first, last; int32_t len, table; };   // offsets into `special_codes` of the two allele passes' strings
struct KmerBuildInputs {
    std::vector<uint8_t> codes, alleles, special_codes;
    std::vector<int64_t> offsets;
    std::vector<KmerSpecialJob> special;
};
bool kmer_build_inputs(const HostIndex &idx, KmerBuildInputs &out);

// nst_nt4_table semantics (libbwa/bntseq.c:38-55): A0 C1 G2 T3, '-' 5, else 4.
const uint8_t *nt4_table();

// ---- fixture-side index construction (the reference's `index` stage is out of
// scope for the hot path; this exists so synthetic benches/tests can run where
// the reference binary is absent).  Bit-compatible with BwtIndexer::BuildIndex.
struct FlankSeq { std::string name; std::string seq; };
bool read_flank_fasta(const std::string &path, std::vector<FlankSeq> &out, std::string &err);
void build_index_from_flanks(const std::vector<FlankSeq> &flanks, bool with_rollhash, HostIndex &out);
bool dump_index(const HostIndex &idx, const std::string &prefix, std::string &err);

}  // namespace fqb
