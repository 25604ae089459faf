// Index loading (reference file formats, read unchanged) and fixture-side
// index construction.  See fq_index.h for the reference anchors.
#include "fq_index.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>

namespace fqb {

static uint8_t *make_nt4() {
    static uint8_t t[256];
    memset(t, 4, sizeof(t));
    t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    t['-'] = 5;
    return t;
}
static uint8_t *g_nt4 = make_nt4();
const uint8_t *nt4_table() { return g_nt4; }

static bool slurp(const std::string &path, std::vector<uint8_t> &out, std::string &err) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) { err = "cannot open " + path; return false; }
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n ? fread(out.data(), 1, (size_t)n, fp) : 0;
    fclose(fp);
    if (got != (size_t)n) { err = "short read on " + path; return false; }
    return true;
}

bool load_bwt(const std::string &bwt_path, const std::string &sa_path, HostBwt &b, std::string &err) {
    std::vector<uint8_t> raw;
    if (!slurp(bwt_path, raw, err)) return false;
    if (raw.size() < 20 || (raw.size() & 3)) { err = "malformed " + bwt_path; return false; }
    const uint32_t *w = reinterpret_cast<const uint32_t *>(raw.data());
    b.primary = w[0];
    b.L2[0] = 0;
    for (int i = 0; i < 4; ++i) b.L2[i + 1] = w[1 + i];
    b.seq_len = b.L2[4];
    b.bwt.assign(w + 5, w + raw.size() / 4);
    if (!slurp(sa_path, raw, err)) return false;
    w = reinterpret_cast<const uint32_t *>(raw.data());
    if (raw.size() < 28 || w[0] != b.primary || w[6] != b.seq_len) {
        err = "SA-BWT inconsistency in " + sa_path; return false;
    }
    b.sa_intv = w[5];
    uint32_t n_sa = (b.seq_len + b.sa_intv) / b.sa_intv;
    if (raw.size() / 4 != 7 + (size_t)n_sa - 1) { err = "unexpected size of " + sa_path; return false; }
    b.sa.resize(n_sa);
    b.sa[0] = 0xffffffffu;
    memcpy(b.sa.data() + 1, w + 7, (size_t)(n_sa - 1) * 4);
    return true;
}

static bool load_ann_amb(const std::string &prefix, HostIndex &idx, std::string &err) {
    std::ifstream ann(prefix + ".ann");
    if (!ann) { err = "cannot open " + prefix + ".ann"; return false; }
    long long l_pac; int n_seqs; unsigned seed;
    ann >> l_pac >> n_seqs >> seed;
    idx.l_pac = l_pac; idx.seed = seed;
    idx.contigs.resize(n_seqs);
    std::string line;
    std::getline(ann, line);
    for (int i = 0; i < n_seqs; ++i) {
        Contig &c = idx.contigs[i];
        std::getline(ann, line);                       // "<gi> <name>[ <anno>]"
        size_t s1 = line.find(' ');
        size_t s2 = line.find(' ', s1 + 1);
        c.name = line.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
        c.anno = s2 == std::string::npos ? "" : line.substr(s2 + 1);
        long long off; int len, nambs;
        ann >> off >> len >> nambs;
        c.offset = off; c.len = len; c.n_ambs = nambs;
        std::getline(ann, line);
    }
    std::ifstream amb(prefix + ".amb");
    if (!amb) { err = "cannot open " + prefix + ".amb"; return false; }
    long long l2; int n2, n_holes;
    amb >> l2 >> n2 >> n_holes;
    if (l2 != l_pac || n2 != n_seqs) { err = "inconsistent .ann and .amb files"; return false; }
    idx.holes.resize(n_holes);
    for (int i = 0; i < n_holes; ++i) {
        long long off; int len; std::string a;
        amb >> off >> len >> a;
        idx.holes[i] = Hole{off, len, a.empty() ? 'N' : a[0]};
    }
    return true;
}

bool load_index(const std::string &prefix, bool rollhash_in_memory, HostIndex &idx, std::string &err) {
    if (!load_bwt(prefix + ".bwt", prefix + ".sa", idx.bwt[0], err)) return false;
    if (!load_bwt(prefix + ".rbwt", prefix + ".rsa", idx.bwt[1], err)) return false;
    if (!load_ann_amb(prefix, idx, err)) return false;
    std::vector<uint8_t> raw;
    if (!slurp(prefix + ".pac", raw, err)) return false;
    size_t need = (size_t)((idx.l_pac + 3) >> 2);
    if (raw.size() < need) { err = ".pac shorter than l_pac"; return false; }
    idx.pac.assign(raw.begin(), raw.begin() + need);
    idx.pac.resize(need + 8, 0);
    if ((uint32_t)idx.l_pac != idx.bwt[0].seq_len || idx.bwt[1].seq_len != idx.bwt[0].seq_len) {
        err = "l_pac and BWT lengths disagree"; return false;
    }
    idx.rollhash_path = prefix + ".rollhash";
    if (rollhash_in_memory) {
        if (!slurp(idx.rollhash_path, idx.rollhash, err)) return false;
        if (idx.rollhash.size() != kRollTableBytes * kNumRollTables) { err = "unexpected .rollhash size"; return false; }
    }
    return true;
}

// ------------------------------------------------------------------------
// construction

bool read_flank_fasta(const std::string &path, std::vector<FlankSeq> &out, std::string &err) {
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    std::string name, seq;
    while (std::getline(in, name)) {                   // two lines per marker, src/BwtIndexer.cpp:868-876
        if (name.empty()) continue;
        if (!std::getline(in, seq)) { err = "odd number of lines in " + path; return false; }
        out.push_back(FlankSeq{name.substr(1), seq});
    }
    return true;
}

// 16-of-32 spaced masks, table order of KmerShrinkage cases 0..5 (src/BwtIndexer.h:262-315)
static inline uint32_t shrink(uint64_t kmer, int which) {
    switch (which) {
    case 0: return (uint32_t)(kmer >> 32);
    case 1: return (uint32_t)kmer;
    case 2: return (uint32_t)((kmer & 0xffff000000000000ull) >> 32) | (uint32_t)(kmer & 0xffff);
    case 3: return (uint32_t)((kmer & 0xffffffff0000ull) >> 16);
    case 4: return (uint32_t)((kmer & 0xffff000000000000ull) >> 32) | (uint32_t)((kmer & 0xffff0000ull) >> 16);
    default: return (uint32_t)((kmer & 0xffff00000000ull) >> 16) | (uint32_t)(kmer & 0xffff);
    }
}

// Same k-mer enumeration as BwtIndexer::AddSeq2HashCore (src/BwtIndexer.cpp:611-713), once per table as AddSeq2Hash
// (src/BwtIndexer.h:317-323) calls it: every 32-mer of the flank, with the centre base replaced by each allele for the
// 32 windows that cover it; after the allele loop the rolling value continues from the LAST allele's window.
// NST_NT4_TABLE (src/BwtIndexer.cpp:59-61) turns every code >= 4 into rand() % 4, drawn afresh at each visit, from
// glibc's never-seeded rand() stream (the caller calls srand(1) so the fixture is reproducible within a process).
static void add_seq_kmers(const std::string &s, const char alleles[2], uint8_t *tables) {
    const size_t n = s.size(), half = n / 2;
    if (n < 32) return;
    auto code = [](char ch) -> uint64_t { const uint8_t c = g_nt4[(uint8_t)ch]; return c < 4 ? c : (uint64_t)(rand() % 4); };
    for (int t = 0; t < kNumRollTables; ++t) {
        auto setbit = [&](uint64_t kmer) {
            uint32_t x = shrink(kmer, t);
            tables[(uint64_t)t * kRollTableBytes + (x >> 3)] |= (uint8_t)(1u << (x & 7));
        };
        uint64_t datum = 0;
        size_t i = 0;
        for (; i < 32; ++i) datum = (datum << 2) | code(s[i]);
        setbit(datum);
        for (; i < half; ++i) { datum = (datum << 2) | code(s[i]); setbit(datum); }
        uint64_t tmp = datum;
        for (int a = 0; a < 2; ++a) {
            tmp = datum;
            for (size_t j = i; j < half + 32 && j < n + 32; ++j) {
                tmp = (tmp << 2) | code(j == half ? alleles[a] : s[j]);
                setbit(tmp);
            }
        }
        datum = tmp;
        for (i = half + 32; i < n; ++i) { datum = (datum << 2) | code(s[i]); setbit(datum); }
    }
}

namespace {
// glibc's rand() (random_r TYPE_3: r[i] = r[i-3] + r[i-31], output >> 1) from its default state srand(1), kept local so the
// library never touches the process-wide generator
struct GlibcRand {
    std::vector<uint32_t> hist;
    GlibcRand() {
        hist.resize(344);
        int32_t x = 1;
        hist[0] = 1;
        for (int i = 1; i < 31; ++i) {
            int64_t v = (16807LL * x) % 2147483647LL;
            if (v < 0) v += 2147483647LL;
            x = (int32_t)v; hist[i] = (uint32_t)x;
        }
        for (int i = 31; i < 34; ++i) hist[i] = hist[i - 31];
        for (int i = 34; i < 344; ++i) hist[i] = hist[i - 31] + hist[i - 3];
    }
    uint32_t next() {
        const size_t k = hist.size();
        hist.push_back(hist[k - 31] + hist[k - 3]);
        return hist.back() >> 1;
    }
};
}  // namespace

bool kmer_build_inputs(const HostIndex &idx, KmerBuildInputs &o) {
    o = KmerBuildInputs();
    for (const Contig &c : idx.contigs) {
        const size_t at = c.name.find('@');
        if (at == std::string::npos || at + 3 >= c.name.size() || c.len < 65) return false;
        o.offsets.push_back(c.offset);
        o.alleles.push_back(g_nt4[(uint8_t)c.name[at + 1]]);
        o.alleles.push_back(g_nt4[(uint8_t)c.name[at + 3]]);
    }
    o.offsets.push_back(idx.l_pac);
    o.codes.resize((size_t)idx.l_pac);
    for (int64_t i = 0; i < idx.l_pac; ++i) o.codes[i] = (idx.pac[i >> 2] >> ((3 - (i & 3)) << 1)) & 3;
    for (const Hole &h : idx.holes)
        for (int64_t i = 0; i < h.len; ++i) o.codes[h.offset + i] = g_nt4[(uint8_t)h.amb];
    GlibcRand rng;
    for (size_t f = 0; f < idx.contigs.size(); ++f) {                    // flank order == the order Fa2Pac hashes them in
        const int64_t off = o.offsets[f];
        const int n = (int)(o.offsets[f + 1] - off), half = n / 2;
        bool special = o.alleles[2 * f] >= 4 || o.alleles[2 * f + 1] >= 4;
        for (int i = 0; i < n && !special; ++i) special = o.codes[off + i] >= 4;
        if (!special) continue;
        o.alleles[2 * f] |= 0x80;
        std::vector<uint8_t> str(n);
        for (int strand = 0; strand < 2; ++strand) {
            for (int i = 0; i < n; ++i) {
                const uint8_t c = strand ? o.codes[off + n - 1 - i] : o.codes[off + i];
                str[i] = strand ? (c < 4 ? (uint8_t)(3 - c) : (uint8_t)4) : c;     // ReverseComplement: anything else -> '\0' -> code 4
            }
            for (int t = 0; t < kNumRollTables; ++t) {
                auto draw = [&](uint8_t c) -> uint8_t { return c < 4 ? c : (uint8_t)(rng.next() % 4); };
                KmerSpecialJob job;
                job.len = n; job.table = t;
                job.first = (int64_t)o.special_codes.size();
                o.special_codes.resize(o.special_codes.size() + 2 * (size_t)n);
                job.last = job.first + n;
                uint8_t *c0 = o.special_codes.data() + job.first, *c1 = o.special_codes.data() + job.last;
                for (int i = 0; i < half; ++i) c0[i] = c1[i] = draw(str[i]);
                for (int a = 0; a < 2; ++a) {
                    uint8_t *ca = a ? c1 : c0;
                    for (int j = half; j < half + 32; ++j) ca[j] = draw(j == half ? (uint8_t)(o.alleles[2 * f + a] & 0x7f) : str[j]);
                }
                for (int i = half + 32; i < n; ++i) c0[i] = c1[i] = draw(str[i]);
                o.special.push_back(job);
            }
        }
    }
    return true;
}

static std::string revcomp_ascii(const std::string &s) {
    std::string r(s.rbegin(), s.rend());
    for (char &c : r) {
        switch (c) {
        case 'A': case 'a': c = 'T'; break;
        case 'C': case 'c': c = 'G'; break;
        case 'G': case 'g': c = 'C'; break;
        case 'T': case 't': c = 'A'; break;
        default: c = 0; break;                         // match_table default-inserts '\0' (src/BwtIndexer.h:236-245)
        }
    }
    return r;
}

// suffix array of a 2-bit text with an implicit smallest sentinel
static void build_bwt(const std::vector<uint8_t> &T, HostBwt &b) {
    const uint32_t n = (uint32_t)T.size();
    const int K = 10;
    const uint32_t nb = 1u << (2 * K);
    std::vector<uint32_t> start(nb + 1, 0), key(n);
    {
        uint32_t v = 0;                                 // key of suffix i = first K symbols, zero padded
        for (int64_t i = (int64_t)n - 1; i >= 0; --i) {
            v = (v >> 2) | ((uint32_t)T[i] << (2 * (K - 1)));
            key[i] = v;
        }
    }
    for (uint32_t i = 0; i < n; ++i) ++start[key[i] + 1];
    for (uint32_t i = 0; i < nb; ++i) start[i + 1] += start[i];
    std::vector<uint32_t> sa(n);
    {
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (uint32_t i = 0; i < n; ++i) sa[fill[key[i]]++] = i;
    }
    key.clear(); key.shrink_to_fit();
    const uint8_t *t = T.data();
    auto less = [t, n](uint32_t a, uint32_t c) {
        uint32_t la = n - a, lc = n - c, m = la < lc ? la : lc;
        int r = memcmp(t + a, t + c, m);
        return r ? r < 0 : la < lc;
    };
    unsigned nth = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nth; ++w)
        pool.emplace_back([&, w]() {
            for (uint32_t bkt = w; bkt < nb; bkt += nth)
                if (start[bkt + 1] - start[bkt] > 1)
                    std::sort(sa.begin() + start[bkt], sa.begin() + start[bkt + 1], less);
        });
    for (auto &th : pool) th.join();

    // rows 0..n: row 0 is the sentinel suffix; the row whose suffix starts at 0 is `primary`
    memset(b.L2, 0, sizeof(b.L2));
    for (uint32_t i = 0; i < n; ++i) ++b.L2[1 + T[i]];
    for (int i = 2; i <= 4; ++i) b.L2[i] += b.L2[i - 1];
    b.seq_len = n;
    std::vector<uint8_t> B(n);
    uint32_t o = 0;
    b.primary = 0;
    B[o++] = T[n - 1];                                  // row 0
    for (uint32_t r = 1; r <= n; ++r) {
        uint32_t p = sa[r - 1];
        if (p == 0) { b.primary = r; continue; }
        B[o++] = T[p - 1];
    }
    // interleave: every 128 bases 4 cumulative counts then 8 words (bwt_bwtupdate_core, src/BwtIndexer.cpp:1369-1392)
    uint32_t n_occ = (n + kOccInterval - 1) / kOccInterval + 1;
    b.bwt.assign((size_t)((n + 15) >> 4) + (size_t)n_occ * 4, 0);
    uint32_t c[4] = {0, 0, 0, 0};
    size_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (i % kOccInterval == 0) { memcpy(&b.bwt[k], c, 16); k += 4; }
        if (i % 16 == 0) ++k;
        b.bwt[k - 1] |= (uint32_t)B[i] << ((15 - (i & 15)) << 1);
        ++c[B[i]];
    }
    memcpy(&b.bwt[k], c, 16);
    // SA samples every 32 rows (bwt_cal_sa, libbwa/bwt.c:48-67)
    b.sa_intv = 32;
    uint32_t n_sa = (n + 32) / 32;
    b.sa.assign(n_sa, 0);
    b.sa[0] = 0xffffffffu;
    for (uint32_t r = 32; r <= n; r += 32) b.sa[r / 32] = sa[r - 1];
}

void build_index_from_flanks(const std::vector<FlankSeq> &flanks, bool with_rollhash, HostIndex &idx) {
    idx = HostIndex();
    idx.seed = 11;                                      // src/BwtIndexer.cpp:849
    srand48(idx.seed);
    srand(1);                                           // the state a fresh process's rand() starts in (see add_seq_kmers)
    std::vector<uint8_t> T;
    if (with_rollhash) idx.rollhash.assign(kRollTableBytes * kNumRollTables, 0);
    for (const FlankSeq &f : flanks) {
        Contig c;
        c.name = f.name; c.anno = "(null)";
        c.len = (int32_t)f.seq.size();
        c.offset = idx.contigs.empty() ? 0 : idx.contigs.back().offset + idx.contigs.back().len;
        if (with_rollhash) {
            size_t at = f.name.find('@');
            char alleles[2] = {f.name[at + 1], f.name[at + 3]};
            add_seq_kmers(f.seq, alleles, idx.rollhash.data());
            add_seq_kmers(revcomp_ascii(f.seq), alleles, idx.rollhash.data());
        }
        int lasts = 0;
        for (size_t i = 0; i < f.seq.size(); ++i) {
            int ch = (uint8_t)f.seq[i];
            int code = g_nt4[ch];
            if (code >= 4) {
                if (lasts == ch && !idx.holes.empty()) ++idx.holes.back().len;
                else { idx.holes.push_back(Hole{c.offset + (int64_t)i, 1, (char)ch}); ++c.n_ambs; }
                code = (int)(lrand48() & 3);
            }
            lasts = ch;
            T.push_back((uint8_t)code);
        }
        idx.contigs.push_back(c);
    }
    idx.l_pac = (int64_t)T.size();
    idx.pac.assign((size_t)((idx.l_pac + 3) >> 2) + 8, 0);
    for (int64_t i = 0; i < idx.l_pac; ++i) idx.pac[i >> 2] |= (uint8_t)(T[i] << ((3 - (i & 3)) << 1));
    std::vector<uint8_t> R(T.rbegin(), T.rend());
    std::thread t1([&]() { build_bwt(T, idx.bwt[0]); });
    build_bwt(R, idx.bwt[1]);
    t1.join();
}

static bool write_file(const std::string &path, const void *p, size_t n, std::string &err, const char *mode = "wb") {
    FILE *fp = fopen(path.c_str(), mode);
    if (!fp) { err = "cannot write " + path; return false; }
    bool ok = n == 0 || fwrite(p, 1, n, fp) == n;
    fclose(fp);
    if (!ok) err = "short write on " + path;
    return ok;
}

bool dump_index(const HostIndex &idx, const std::string &prefix, std::string &err) {
    for (int s = 0; s < 2; ++s) {
        const HostBwt &b = idx.bwt[s];
        std::vector<uint32_t> out;
        out.push_back(b.primary);
        for (int i = 1; i <= 4; ++i) out.push_back(b.L2[i]);
        out.insert(out.end(), b.bwt.begin(), b.bwt.end());
        if (!write_file(prefix + (s ? ".rbwt" : ".bwt"), out.data(), out.size() * 4, err)) return false;
        out.resize(5);
        out.push_back(b.sa_intv);
        out.push_back(b.seq_len);
        out.insert(out.end(), b.sa.begin() + 1, b.sa.end());
        if (!write_file(prefix + (s ? ".rsa" : ".sa"), out.data(), out.size() * 4, err)) return false;
    }
    {   // .pac with its trailer (src/BwtIndexer.cpp:962-975), .rpac (Fa2RevPac, :1285-1310)
        size_t nbytes = (size_t)(idx.l_pac >> 2) + ((idx.l_pac & 3) ? 1 : 0);
        std::vector<uint8_t> out(idx.pac.begin(), idx.pac.begin() + nbytes);
        if (idx.l_pac % 4 == 0) out.push_back(0);
        out.push_back((uint8_t)(idx.l_pac % 4));
        if (!write_file(prefix + ".pac", out.data(), out.size(), err)) return false;
        std::vector<uint8_t> r((size_t)(idx.l_pac >> 2) + 1, 0);
        for (int64_t i = idx.l_pac - 1, j = 0; i >= 0; --i, ++j) {
            int c = idx.pac[i >> 2] >> ((~i & 3) << 1) & 3;
            r[j >> 2] |= (uint8_t)(c << ((~j & 3) << 1));
        }
        r.push_back((uint8_t)(idx.l_pac % 4));
        if (!write_file(prefix + ".rpac", r.data(), r.size(), err)) return false;
    }
    {
        std::string ann, amb;
        char buf[512];
        snprintf(buf, sizeof buf, "%lld %d %u\n", (long long)idx.l_pac, (int)idx.contigs.size(), idx.seed);
        ann += buf;
        for (const Contig &c : idx.contigs) {
            ann += "0 " + c.name;
            if (!c.anno.empty()) ann += " " + c.anno;
            ann += "\n";
            snprintf(buf, sizeof buf, "%lld %d %d\n", (long long)c.offset, c.len, c.n_ambs);
            ann += buf;
        }
        snprintf(buf, sizeof buf, "%lld %d %u\n", (long long)idx.l_pac, (int)idx.contigs.size(), (unsigned)idx.holes.size());
        amb += buf;
        for (const Hole &h : idx.holes) {
            snprintf(buf, sizeof buf, "%lld %d %c\n", (long long)h.offset, h.len, h.amb);
            amb += buf;
        }
        if (!write_file(prefix + ".ann", ann.data(), ann.size(), err)) return false;
        if (!write_file(prefix + ".amb", amb.data(), amb.size(), err)) return false;
    }
    if (!idx.rollhash.empty())
        if (!write_file(prefix + ".rollhash", idx.rollhash.data(), idx.rollhash.size(), err)) return false;
    return true;
}

}  // namespace fqb
