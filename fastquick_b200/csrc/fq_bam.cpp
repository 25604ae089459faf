#include "fq_bam.h"

#include "../../include/fastquick_b200.h"
#include "fq_common.h"
#include "fq_deflate.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <thread>

namespace fqb {

// ---------------------------------------------------------------------------------------------- BGZF
namespace {
constexpr size_t kBlockIn = 0xff00;       // payload bytes per block (htslib's BGZF_BLOCK_SIZE)

// The BAM members are compressed by this library's one-shot deflate (fq_deflate.cpp: 2-2.5x the speed of zlib's level 1
// for members ~4 % smaller on FASTQ-like payload); FQB_BAM_LEVEL=0..9 asks for zlib at that level instead.  The records,
// not the compressed bytes, are what is compared with the reference's file.
int bgzf_level() {
    static const int lvl = []() { const char *e = getenv("FQB_BAM_LEVEL"); const int v = e ? atoi(e) : -1; return v < 0 || v > 9 ? -1 : v; }();
    return lvl;
}

constexpr size_t kSlot = 18 + kBlockIn + 1024 + 8;      // room for any member: compressBound(0xff00) = 0xff00 + 33 with these settings

// one BGZF member for data[0..n), n <= kBlockIn, written to out[0..kSlot); returns its size
size_t bgzf_block(const char *data, size_t n, char *out) {
    const int level = bgzf_level();
    size_t clen;
    if (level < 0) {
        clen = deflate_fast((const uint8_t *)data, n, (uint8_t *)out + 18, kSlot - 18 - 8);
    } else {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = (Bytef *)data; zs.avail_in = (uInt)n;
        zs.next_out = (Bytef *)out + 18; zs.avail_out = (uInt)(kSlot - 18 - 8);
        deflate(&zs, Z_FINISH);
        clen = zs.total_out;
        deflateEnd(&zs);
    }
    const size_t total = 18 + clen + 8;
    static const unsigned char hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    out[16] = (char)((total - 1) & 0xff); out[17] = (char)((total - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)data, (uInt)n), isz = (uint32_t)n;
    memcpy(out + 18 + clen, &crc, 4);
    memcpy(out + 18 + clen + 4, &isz, 4);
    return total;
}
}  // namespace

constexpr size_t kChunkBlocks = 64;       // members handed to the writer thread at a time (4 MiB of payload)
constexpr size_t kMaxQueued = 64;          // chunks waiting for the writer thread before write() blocks (256 MiB)

bool BgzfWriter::open(const std::string &path, std::string &err) {
    fp_ = fopen(path.c_str(), "wb");
    if (!fp_) { err = "cannot write " + path; return false; }
    pending_.clear(); failed_ = false; closing_ = false; queue_.clear();
    writer_ = std::thread([this]() { run(); });
    return true;
}
void BgzfWriter::write(const void *data, size_t n) {
    pending_.append((const char *)data, n);
    if (pending_.size() >= kChunkBlocks * kBlockIn) hand_over(false);
}
void BgzfWriter::write_owned(std::string &&data) {
    hand_over(true);                                // members may end anywhere: BAM records are free to straddle them
    if (data.empty()) return;
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&]() { return queue_.size() < kMaxQueued; });
        queue_.push_back(std::move(data));
    }
    cv_.notify_all();
}
void BgzfWriter::hand_over(bool all) {
    const size_t whole = all ? pending_.size() : pending_.size() / kBlockIn * kBlockIn;
    if (!whole) return;
    std::string chunk;
    if (whole == pending_.size()) chunk.swap(pending_);
    else { chunk.assign(pending_, 0, whole); pending_.erase(0, whole); }
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&]() { return queue_.size() < kMaxQueued; });
        queue_.push_back(std::move(chunk));
    }
    cv_.notify_all();
}
// The writer thread and its helpers: the helpers live as long as the file is open (a thread per chunk would cost more than
// a chunk's members now that compressing one takes a few hundred microseconds) and take members off a shared counter.
struct BgzfWriter::Crew {
    std::mutex m; std::condition_variable start, done;
    uint64_t generation = 0; unsigned busy = 0; bool stop = false;
    const std::string *data = nullptr; std::atomic<size_t> next{0};
    std::vector<char> slots; std::vector<size_t> sizes;          // member b of the current chunk: slots[b * kSlot .. + sizes[b])
    std::vector<std::thread> helpers;
    void members() {
        const size_t n_blocks = sizes.size();
        for (size_t b; (b = next.fetch_add(1)) < n_blocks;) {
            const size_t off = b * kBlockIn, len = std::min(kBlockIn, data->size() - off);
            sizes[b] = bgzf_block(data->data() + off, len, slots.data() + b * kSlot);
        }
    }
    void helper() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m);
                start.wait(l, [&]() { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            members();
            { std::lock_guard<std::mutex> l(m); --busy; }
            done.notify_all();
        }
    }
    void compress(const std::string &d) {
        data = &d; next.store(0);
        sizes.assign((d.size() + kBlockIn - 1) / kBlockIn, 0);
        if (slots.size() < sizes.size() * kSlot) slots.resize(sizes.size() * kSlot);
        { std::lock_guard<std::mutex> l(m); ++generation; busy = (unsigned)helpers.size(); }
        start.notify_all();
        members();
        std::unique_lock<std::mutex> l(m);
        done.wait(l, [&]() { return busy == 0; });
    }
};

void BgzfWriter::run() {
    Crew crew;
    unsigned nthr = std::thread::hardware_concurrency();
    if (nthr < 1) nthr = 1;
    if (nthr > 16) nthr = 16;
    for (unsigned t = 1; t < nthr; ++t) crew.helpers.emplace_back([&crew]() { crew.helper(); });
    for (;;) {
        std::string chunk;
        {
            std::unique_lock<std::mutex> l(m_);
            cv_.wait(l, [&]() { return closing_ || !queue_.empty(); });
            if (queue_.empty()) break;
            chunk.swap(queue_.front()); queue_.pop_front();
        }
        cv_.notify_all();
        crew.compress(chunk);
        for (size_t b = 0; b < crew.sizes.size(); ++b)
            if (fwrite(crew.slots.data() + b * kSlot, 1, crew.sizes[b], fp_) != crew.sizes[b]) failed_ = true;
    }
    { std::lock_guard<std::mutex> l(crew.m); crew.stop = true; }
    crew.start.notify_all();
    for (auto &t : crew.helpers) t.join();
}
bool BgzfWriter::close(std::string &err) {
    if (!fp_) return true;
    hand_over(true);
    { std::lock_guard<std::mutex> l(m_); closing_ = true; }
    cv_.notify_all();
    if (writer_.joinable()) writer_.join();
    static const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fwrite(eof, 1, 28, fp_) != 28) failed_ = true;
    if (fclose(fp_) != 0) failed_ = true;
    fp_ = nullptr;
    if (failed_) { err = "write error on the BAM file"; return false; }
    return true;
}
BgzfWriter::~BgzfWriter() { std::string e; close(e); }

// ---------------------------------------------------------------------------------------------- header
static void put32(std::string &o, int32_t v) { o.append((const char *)&v, 4); }

bool bam_prepare(const HostIndex &idx, const fqb_gap_opt_t &g, const std::vector<std::pair<std::string, int>> &genome_contigs,
                 const std::string &rg_line, BamContext &ctx, std::string &hb, std::string &err) {
    ctx.idx = &idx; ctx.gopt = g; ctx.refs = genome_contigs; ctx.rg_id.clear();
    // SetSamFileHeader: @PG, then the @RG tags of --RG, then one @SQ per line of <reference>.fai
    std::string text = "@PG\tID:FASTQuick\tVN:0.0.1\n";
    if (rg_line.compare(0, 3, "@RG") == 0) {
        // bwa_set_rg (libbwa): the ID value becomes bwa_rg_id; SetSamFileHeader re-emits every TAG:value token
        std::string line = rg_line;
        for (size_t k = 0; k + 1 < line.size(); ++k)
            if (line[k] == '\\' && line[k + 1] == 't') { line[k] = '\t'; line.erase(k + 1, 1); }
        std::stringstream ss(line);
        std::string tok, out = "@RG";
        while (ss >> tok) {
            if (tok == "@RG" || tok.size() < 3) continue;
            if (tok.compare(0, 3, "ID:") == 0) ctx.rg_id = tok.substr(3);
        }
        // libStatGen prints ID first, then the other tags in the order they were set
        std::stringstream s2(line);
        if (!ctx.rg_id.empty()) out += "\tID:" + ctx.rg_id;
        while (s2 >> tok) {
            if (tok == "@RG" || tok.size() < 3 || tok.compare(0, 3, "ID:") == 0) continue;
            out += "\t" + tok;
        }
        text += out + "\n";
    }
    std::map<std::string, int> ref_id;
    for (size_t i = 0; i < ctx.refs.size(); ++i) {
        text += "@SQ\tSN:" + ctx.refs[i].first + "\tLN:" + std::to_string(ctx.refs[i].second) + "\n";
        ref_id[ctx.refs[i].first] = (int)i;
    }
    hb.assign("BAM\1", 4);
    put32(hb, (int32_t)text.size());
    hb += text;
    put32(hb, (int32_t)ctx.refs.size());
    for (auto &r : ctx.refs) { put32(hb, (int32_t)r.first.size() + 1); hb.append(r.first.c_str(), r.first.size() + 1); put32(hb, r.second); }
    // genome coordinates of the flanks: "<chrom>:<pos>@<ref>/<alt>[L]" (SetSamRecord's BAM_DEBUG branch)
    const size_t nc = idx.contigs.size();
    ctx.ref_of_contig.assign(nc, -1); ctx.ref_coord.assign(nc, 0); ctx.is_long.assign(nc, 0); ctx.chrom_of_contig.assign(nc, "");
    for (size_t c = 0; c < nc; ++c) {
        const std::string &nm = idx.contigs[c].name;
        const size_t colon = nm.find(':');
        ctx.chrom_of_contig[c] = nm.substr(0, colon);
        if (colon != std::string::npos) ctx.ref_coord[c] = (int)strtol(nm.c_str() + colon + 1, nullptr, 10);
        ctx.is_long[c] = !nm.empty() && nm.back() == 'L';
        auto it = ref_id.find(ctx.chrom_of_contig[c]);
        if (it == ref_id.end()) { err = "chromosome " + ctx.chrom_of_contig[c] + " of flank " + nm + " is not in the reference .fai"; return false; }
        ctx.ref_of_contig[c] = it->second;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------- records
namespace {
enum { FPD = 1, FPP = 2, FSU = 4, FMU = 8, FSR = 16, FMR = 32 };
constexpr int kNoMatch = 0, kMateSW = 3;

struct Coor { int seqid; int nn; };
// bns_coor_pac2real (libbwa/bntseq.c:268-303)
Coor pac2real(const HostIndex &I, int64_t pac, int len) {
    int left = 0, mid = 0, right = (int)I.contigs.size();
    while (left < right) {
        mid = (left + right) >> 1;
        if (pac >= I.contigs[mid].offset) {
            if (mid == (int)I.contigs.size() - 1) break;
            if (pac < I.contigs[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    Coor c; c.seqid = mid; c.nn = 0;
    int l = 0, r = (int)I.holes.size();
    while (l < r) {
        const int m = (l + r) >> 1;
        const Hole &h = I.holes[m];
        if (pac >= h.offset + h.len) l = m + 1;
        else if (pac + len <= h.offset) r = m;
        else {
            if (pac >= h.offset) c.nn += h.offset + h.len < pac + len ? (int)(h.offset + h.len - pac) : len;
            else c.nn += h.offset + h.len < pac + len ? h.len : (int)(len - (h.offset - pac));
            break;
        }
    }
    return c;
}
int64_t pos_end(const fqb_read_t &p) {
    if (p.has_cigar) {
        int64_t x = p.pos;
        for (int j = 0; j < p.n_cigar; ++j) { const int op = p.cigar[j] >> 14; if (op == 0 || op == 2) x += p.cigar[j] & 0x3fff; }
        return x;
    }
    return (int64_t)p.pos + p.len;
}
int64_t pos_5(const fqb_read_t &p) { return p.type != kNoMatch ? (p.strand ? pos_end(p) : (int64_t)p.pos) : -1; }
int real_start(const BamContext &C, int seqid, int64_t pac_pos) {
    const int pos = (int)(pac_pos - C.idx->contigs[seqid].offset + 1);
    return C.ref_coord[seqid] - (C.is_long[seqid] ? C.gopt.flank_long_len : C.gopt.flank_len) + pos - 1;
}
int reg2bin(int32_t beg, int32_t end) {
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
}
void tag_int(std::string &o, const char *t, int v) {        // SamRecord::addIntTag's choice of the BAM integer type
    o.append(t, 2);
    if (v < 0) {
        if (v > -128) { o.push_back('c'); o.push_back((char)(int8_t)v); }
        else if (v > -32768) { o.push_back('s'); int16_t x = (int16_t)v; o.append((const char *)&x, 2); }
        else { o.push_back('i'); o.append((const char *)&v, 4); }
    } else {
        if (v < 255) { o.push_back('C'); o.push_back((char)(uint8_t)v); }
        else if (v < 65535) { o.push_back('S'); uint16_t x = (uint16_t)v; o.append((const char *)&x, 2); }
        else { o.push_back('I'); uint32_t x = (uint32_t)v; o.append((const char *)&x, 4); }
    }
}
void tag_str(std::string &o, const char *t, const std::string &v) { o.append(t, 2); o.push_back('Z'); o.append(v.c_str(), v.size() + 1); }
void put_num(std::string &o, long long v) {
    char b[24]; int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) b[n++] = '-';
    while (n) o.push_back(b[--n]);
}
// per-thread scratch of the formatter (records are formatted by a few host threads, one batch slice each)
struct Scratch { std::string tags, md, xa; std::vector<uint8_t> sq; };
thread_local Scratch g_scratch;

// the read in alignment orientation as nt4 codes
void oriented(const uint8_t *bases, int len, int strand, std::vector<uint8_t> &out) {
    const uint8_t *t = nt4_table();
    out.resize((size_t)len);
    for (int j = 0; j < len; ++j) {
        if (!strand) { const uint8_t c = t[bases[j]]; out[j] = c > 4 ? 4 : c; }
        else { const uint8_t c = t[bases[len - 1 - j]]; out[j] = c < 4 ? (uint8_t)(3 - c) : 4; }
    }
}
// MD tag of bwa_cal_md1 (libbwa/bwase.c:234-296); seq = the trimmed read in alignment orientation
void md_string(const HostIndex &I, const fqb_read_t &s, const uint16_t *cigar, int n_cigar, bool has_cigar, int len, const uint8_t *seq, std::string &o) {
    o.clear();
    auto base = [&](int64_t k) { return (I.pac[(size_t)(k >> 2)] >> ((~k & 3) << 1)) & 3; };
    int u = 0;
    int64_t x = s.pos, y = 0;
    if (has_cigar) {
        for (int k = 0; k < n_cigar; ++k) {
            const int l = cigar[k] & 0x3fff, op = cigar[k] >> 14;
            if (op == 0) {
                for (int z = 0; z < l && x + z < I.l_pac; ++z) {
                    const int c = base(x + z);
                    if (seq[y + z] > 3 || c != seq[y + z]) { put_num(o, u); o.push_back("ACGTN"[c]); u = 0; } else ++u;
                }
                x += l; y += l;
            } else if (op == 1 || op == 3) y += l;
            else {
                put_num(o, u); o.push_back('^');
                for (int z = 0; z < l && x + z < I.l_pac; ++z) o.push_back("ACGT"[base(x + z)]);
                u = 0; x += l;
            }
        }
    } else {
        for (int z = 0; z < len; ++z) {
            const int c = base(x + z);
            if (seq[y + z] > 3 || c != seq[y + z]) { put_num(o, u); o.push_back("ACGTN"[c]); u = 0; } else ++u;
        }
    }
    put_num(o, u);
}

// mate_ptr == nullptr: single-end input (SetSamRecord(..., mate = 0, ...))
// rseq: what the reference's p->rseq buffer holds for this read slot (nt4 codes, full_len bytes): the reverse complement of
// the trimmed read followed by whatever an earlier occupant of the slot left there (nullptr = zeros, the single-end reader
// callocs the buffer per read).  Only SetSamRecord's "no match" branch looks at it.
void one_record(const BamContext &C, fqb_read_t &p, const fqb_read_t *mate_ptr, const char *name, const uint8_t *bases, const uint8_t *quals,
                const XaHit *xa, int n_xa, const uint8_t *rseq, std::string &out) {
    fqb_read_t absent;
    memset(&absent, 0, sizeof absent);
    const bool has_mate = mate_ptr != nullptr;
    const fqb_read_t &mate = has_mate ? *mate_ptr : absent;
    const HostIndex &I = *C.idx;
    int flag = p.extra_flag, j, am = 0;
    if (p.type == kNoMatch && mate.type == kNoMatch) {
        // SetSamRecord's "this read has no match" branch (1227-1259): reachable when the bridge check demoted the pair
        flag |= FSU;
        if (has_mate) flag |= FMU;
        const int L = p.len;
        const uint8_t *t = nt4_table();
        std::string seq((size_t)L, 'N'), qual((size_t)L, '\0'), tags;
        for (int k = 0; k < L; ++k) {
            if (!p.strand) { const uint8_t c = t[bases[k]]; seq[k] = "ACGTN"[c > 4 ? 4 : c]; }
            else if (rseq) { const uint8_t c = rseq[k]; seq[k] = "ACGTN"[c > 4 ? 4 : c]; }       // s = p->rseq, printed over the FULL length
            else if (k < p.clip_len) { const uint8_t c = t[bases[p.clip_len - 1 - k]]; seq[k] = "TGCAN"[c > 4 ? 4 : c]; }
            else seq[k] = 'A';                                                                   // calloc'ed tail of the single-end reader
            qual[k] = (char)((p.strand ? quals[L - 1 - k] : quals[k]) - 33);
        }
        if (!C.rg_id.empty()) tag_str(tags, "RG", C.rg_id);
        if (p.clip_len < p.full_len) tag_int(tags, "XC", p.clip_len);
        const int l_name = (int)strlen(name) + 1;
        put32(out, 32 + l_name + (L + 1) / 2 + L + (int)tags.size());
        put32(out, -1); put32(out, -1);
        out.push_back((char)l_name); out.push_back((char)0);
        { uint16_t b = (uint16_t)reg2bin(-1, 0); out.append((const char *)&b, 2); }
        { uint16_t n = 0; out.append((const char *)&n, 2); }
        { uint16_t f = (uint16_t)flag; out.append((const char *)&f, 2); }
        put32(out, L); put32(out, -1); put32(out, -1); put32(out, 0);
        out.append(name, (size_t)l_name);
        for (int k = 0; k < L; k += 2) {
            static const char *codes = "=ACMGRSVTWYHKDBN";
            const int hi = (int)(strchr(codes, seq[k]) - codes), lo = k + 1 < L ? (int)(strchr(codes, seq[k + 1]) - codes) : 0;
            out.push_back((char)(hi << 4 | lo));
        }
        out += qual; out += tags;
        return;
    }
    if (p.type == kNoMatch) { p.pos = mate.pos; p.strand = mate.strand; flag |= FSU; j = 1; }
    else j = (int)(pos_end(p) - p.pos);
    const Coor co = pac2real(I, p.pos, j);
    const int seqid = co.seqid;
    int nn = co.nn;
    if (p.type != kNoMatch && (int64_t)p.pos + j - I.contigs[seqid].offset > I.contigs[seqid].len) flag |= FSU;
    if (p.strand) flag |= FSR;
    if (has_mate) { if (mate.type != kNoMatch) { if (mate.strand) flag |= FMR; } else flag |= FMU; }
    int ref_id = -1, pos1 = 0, read_real_start = 0;
    if (p.type != kNoMatch) { read_real_start = real_start(C, seqid, p.pos); ref_id = C.ref_of_contig[seqid]; pos1 = read_real_start; }
    // CIGAR
    uint32_t cig[FQB_MAX_CIGAR + 1]; int n_cig = 0;
    if (p.type != kNoMatch) {
        if (p.has_cigar) for (int k = 0; k < p.n_cigar; ++k) { static const int opmap[4] = {0, 1, 2, 4}; cig[n_cig++] = (uint32_t)(p.cigar[k] & 0x3fff) << 4 | opmap[p.cigar[k] >> 14]; }
        else cig[n_cig++] = (uint32_t)p.len << 4;
    }
    // mate fields
    int mref = -1, mpos1 = 0;
    long long isize = 0;
    if (mate.type != kNoMatch) {
        am = mate.seQ < p.seQ ? mate.seQ : p.seQ;
        const Coor mc = pac2real(I, mate.pos, mate.len);
        const int m_start = real_start(C, mc.seqid, mate.pos);
        mref = seqid == mc.seqid ? ref_id : C.ref_of_contig[mc.seqid];       // "=" resolves to this record's own reference name
        isize = seqid == mc.seqid ? pos_5(mate) - pos_5(p) : 0;
        if (p.type == kNoMatch) isize = 0;
        mpos1 = m_start;
        read_real_start = m_start;                // the reference reuses the variable; only read again when the mate is unmapped
    } else if (has_mate) { mref = ref_id; mpos1 = read_real_start; isize = 0; }
    else { mref = -1; mpos1 = 0; isize = 0; }
    // sequence and qualities in alignment orientation (full length) are written straight into the record below
    const int L = p.full_len;
    const uint8_t *t = nt4_table();
    // tags
    std::string &tags = g_scratch.tags;
    tags.clear();
    if (!C.rg_id.empty()) tag_str(tags, "RG", C.rg_id);
    if (p.clip_len < p.full_len) tag_int(tags, "XC", p.clip_len);
    if (p.type != kNoMatch) {
        char xt = "NURM"[p.type & 3];
        if (nn > 10) xt = 'N';
        tags.append("XTA", 3); tags.push_back(xt);
        tag_int(tags, (C.gopt.mode & 0x02) ? "NM" : "CM", p.nm);          // BWA_MODE_COMPREAD
        if (nn) tag_int(tags, "XN", nn);
        if (has_mate) { tag_int(tags, "SM", p.seQ); tag_int(tags, "AM", am); }
        if (p.type != kMateSW) {
            tag_int(tags, "X0", (int)p.c1);
            if ((int)p.c1 <= C.gopt.max_top2) tag_int(tags, "X1", (int)p.c2);
        }
        tag_int(tags, "XM", p.n_mm);
        tag_int(tags, "XO", p.n_gapo);
        tag_int(tags, "XG", p.n_gapo + p.n_gape);
        // MD over the trimmed read: strip the soft clip bwa_correct_trimmed appended
        {
            std::vector<uint8_t> &sq = g_scratch.sq;
            oriented(bases, L, p.strand, sq);
            md_string(I, p, p.cigar, p.n_cigar, p.has_cigar != 0, p.len, sq.data(), g_scratch.md);
            tag_str(tags, "MD", g_scratch.md);
        }
        if (n_xa) {
            std::string &s = g_scratch.xa;
            s.clear();
            for (int i = 0; i < n_xa; ++i) {
                const XaHit &q = xa[i];
                int64_t e = q.pos;
                if (q.has_cigar) { for (int k = 0; k < q.n_cigar; ++k) { const int op = q.cigar[k] >> 14; if (op == 0 || op == 2) e += q.cigar[k] & 0x3fff; } }
                else e += p.len;
                const Coor qc = pac2real(I, q.pos, (int)(e - q.pos));
                s += I.contigs[qc.seqid].name; s.push_back(',');
                s.push_back(q.strand ? '-' : '+');
                put_num(s, (long long)q.pos - I.contigs[qc.seqid].offset + 1); s.push_back(',');
                if (q.has_cigar) for (int k = 0; k < q.n_cigar; ++k) { put_num(s, q.cigar[k] & 0x3fff); s.push_back("MIDS"[q.cigar[k] >> 14]); }
                else { put_num(s, p.len); s.push_back('M'); }
                s.push_back(','); put_num(s, q.gap + q.mm); s.push_back(';');
            }
            tag_str(tags, "XA", s);
        }
    }
    // assemble
    const int l_name = (int)strlen(name) + 1;
    const int32_t pos0 = pos1 - 1, aln_len = p.type != kNoMatch ? (int32_t)(pos_end(p) - p.pos) : 0;
    const int32_t end1 = aln_len ? pos0 + aln_len : pos0 + 1;
    const int bin = reg2bin(pos0, end1);
    const int32_t block = 32 + l_name + 4 * n_cig + (L + 1) / 2 + L + (int)tags.size();
    const size_t at = out.size();
    out.resize(at + 4 + (size_t)block);
    uint8_t *w = (uint8_t *)&out[at];
    auto w32 = [&](int32_t v) { memcpy(w, &v, 4); w += 4; };
    auto w16 = [&](uint16_t v) { memcpy(w, &v, 2); w += 2; };
    w32(block); w32(ref_id); w32(pos0);
    *w++ = (uint8_t)l_name; *w++ = (uint8_t)p.mapQ;
    w16((uint16_t)bin); w16((uint16_t)n_cig); w16((uint16_t)flag);
    w32(L); w32(mref); w32(mpos1 - 1); w32((int32_t)isize);
    memcpy(w, name, (size_t)l_name); w += l_name;
    memcpy(w, cig, 4 * (size_t)n_cig); w += 4 * n_cig;
    // 4-bit bases "=ACMGRSVTWYHKDBN": A 1, C 2, G 4, T 8, N 15; the reverse strand prints the complement of the reversed read
    static const uint8_t nib_f[5] = {1, 2, 4, 8, 15}, nib_r[5] = {8, 4, 2, 1, 15};
    auto nib = [&](int k) -> uint8_t {
        if (k >= L) return 0;
        const uint8_t c = p.strand ? t[bases[L - 1 - k]] : t[bases[k]];
        return p.strand ? nib_r[c > 4 ? 4 : c] : nib_f[c > 4 ? 4 : c];
    };
    for (int k = 0; k < L; k += 2) *w++ = (uint8_t)(nib(k) << 4 | nib(k + 1));
    if (p.strand) for (int k = 0; k < L; ++k) *w++ = (uint8_t)(quals[L - 1 - k] - 33);
    else for (int k = 0; k < L; ++k) *w++ = (uint8_t)(quals[k] - 33);
    memcpy(w, tags.data(), tags.size());
}
}  // namespace

void bam_append_pair(const BamContext &C, fqb_read_t p, fqb_read_t q, const char *name, const char *name_q, const uint8_t *bases_p, const uint8_t *quals_p,
                     const uint8_t *bases_q, const uint8_t *quals_q, const XaHit *xa_p, int n_xa_p, const XaHit *xa_q, int n_xa_q,
                     const uint8_t *rseq_p, const uint8_t *rseq_q, std::string &out) {
    one_record(C, p, &q, name, bases_p, quals_p, xa_p, n_xa_p, rseq_p, out);       // may rewrite p's pos/strand (unmapped read of a half-mapped pair)
    one_record(C, q, &p, name_q, bases_q, quals_q, xa_q, n_xa_q, rseq_q, out);
}

void bam_append_single(const BamContext &C, fqb_read_t p, const char *name, const uint8_t *bases, const uint8_t *quals, const XaHit *xa, int n_xa,
                       std::string &out) {
    one_record(C, p, nullptr, name, bases, quals, xa, n_xa, nullptr, out);
}

}  // namespace fqb

// The BAM writer's member format on a buffer (test hook): data[0..n) cut into 0xff00-byte payloads, each compressed into
// one BGZF member exactly as BgzfWriter does, members back to back in out, no end-of-file member.
extern "C" int fqb_bgzf_compress(const uint8_t *data, int64_t n, uint8_t *out, int64_t cap, int64_t *n_out) {
    if ((!data && n > 0) || n < 0 || (!out && cap > 0) || cap < 0 || !n_out) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    int64_t o = 0;
    for (int64_t off = 0; off < n; off += (int64_t)fqb::kBlockIn) {
        std::vector<char> m(fqb::kSlot);
        const size_t len = fqb::bgzf_block((const char *)data + off, (size_t)std::min<int64_t>((int64_t)fqb::kBlockIn, n - off), m.data());
        if (o + (int64_t)len > cap) { fqb::set_error("output buffer too small"); return FQB_ERR_ARG; }
        memcpy(out + o, m.data(), len);
        o += (int64_t)len;
    }
    *n_out = o;
    return FQB_OK;
}

// BgzfWriter itself on a buffer (test hook): data[0..n) handed to the writer in pieces of `piece` bytes, by write() when
// owned == 0 and by write_owned() otherwise, into a BGZF file with its end-of-file member.
extern "C" int fqb_bgzf_write_file(const char *path, const uint8_t *data, int64_t n, int64_t piece, int32_t owned) {
    if (!path || (!data && n > 0) || n < 0 || piece < 1) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    fqb::BgzfWriter w;
    std::string err;
    if (!w.open(path, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    for (int64_t off = 0; off < n; off += piece) {
        const size_t len = (size_t)std::min<int64_t>(piece, n - off);
        if (owned) w.write_owned(std::string((const char *)data + off, len));
        else w.write(data + off, len);
    }
    if (!w.close(err)) { fqb::set_error(err); return FQB_ERR_IO; }
    return FQB_OK;
}
