// BAM emission for the align stage: the records BwtMapper::SetSamRecord builds
// (src/BwtMapper.cpp:977-1264, compiled with BAM_DEBUG: genome coordinates recovered from the flank
// names) and the header of SetSamFileHeader (947-975), written as BGZF by host threads.
// Input: the per-read result rows computed on the GPU, the multi-hit list kernel's output, and the
// host copies of the reads.  Everything here is formatting; no alignment decision is taken on the host.
#pragma once
#include <cstdint>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/fastquick_b200.h"
#include "fq_index.h"

namespace fqb {

// BGZF: independent <= 64 KiB gzip members.  write() only queues the bytes; a writer thread compresses each queued chunk
// with a few host threads and writes the members in order, so compression overlaps the next batch's GPU work.
class BgzfWriter {
public:
    bool open(const std::string &path, std::string &err);
    void write(const void *data, size_t n);
    void write_owned(std::string &&data);           // whole records; becomes its own run of members, no copy
    bool close(std::string &err);                   // drains the queue, writes the EOF block
    ~BgzfWriter();
private:
    void hand_over(bool all);
    void run();
    struct Crew;
    FILE *fp_ = nullptr;
    std::string pending_;
    bool failed_ = false;
    std::thread writer_;
    std::mutex m_; std::condition_variable cv_;
    std::deque<std::string> queue_;
    bool closing_ = false;
};

struct XaHit { uint32_t pos; uint8_t strand, gap, mm, has_cigar, n_cigar; uint16_t cigar[FQB_MAX_CIGAR]; };

struct BamContext {
    const HostIndex *idx = nullptr;
    fqb_gap_opt_t gopt;
    std::string rg_id;                              // bwa_rg_id ("" = no RG tag)
    std::vector<std::pair<std::string, int>> refs;  // @SQ: BwtIndexer::contigSize (<reference>.fai)
    std::vector<int> ref_of_contig;                 // flank contig -> index into refs (-1 = chromosome not in the .fai)
    std::vector<int> ref_coord;                     // marker coordinate parsed from the flank name
    std::vector<uint8_t> is_long;                   // flank name ends in 'L'
    std::vector<std::string> chrom_of_contig;
};

// "@PG ... @RG ... @SQ ..." text + binary reference list; rg_line as passed to --RG (or empty)
bool bam_prepare(const HostIndex &idx, const fqb_gap_opt_t &g, const std::vector<std::pair<std::string, int>> &genome_contigs,
                 const std::string &rg_line, BamContext &ctx, std::string &header_bytes, std::string &err);

// Appends the two records of one pair (nothing when both reads are filtered or both unmapped; the caller decides
// that with the types the reads had BEFORE StatCollector::AddAlignment's bridge check demoted any of them, as
// the reference's loop does, src/BwtMapper.cpp:2058-2066).
// p, q: result rows (taken by value: SetSamRecord edits the unmapped read of a half-mapped pair).
// bases/quals: the reads as they came from the FASTQ files (ASCII), full_len bytes each.
void bam_append_pair(const BamContext &ctx, fqb_read_t p, fqb_read_t q, const char *name, const char *name_q, const uint8_t *bases_p, const uint8_t *quals_p,
                     const uint8_t *bases_q, const uint8_t *quals_q, const XaHit *xa_p, int n_xa_p, const XaHit *xa_q, int n_xa_q,
                     const uint8_t *rseq_p, const uint8_t *rseq_q, std::string &out);
// rseq_p / rseq_q: the slot's p->rseq buffer as the reference's paired reader leaves it (see fqb_bam_emit), or nullptr

// Single-end input: the record SetSamRecord(bns, p, 0, ...) builds (SingleEndMapper, src/BwtMapper.cpp:1384-1388)
void bam_append_single(const BamContext &ctx, fqb_read_t p, const char *name, const uint8_t *bases, const uint8_t *quals, const XaHit *xa, int n_xa,
                       std::string &out);

}  // namespace fqb
