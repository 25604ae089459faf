// Synthetic reduced reference and read simulator.  Integer-only randomness
// (splitmix64 + Irwin-Hall normals) so outputs are identical on every host.
#include "fq_synth.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>

namespace fqb {

namespace {
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    // uniform in [0,1) with 24 bits -- compared against rates only
    double unif() { return (double)(next() >> 40) * (1.0 / 16777216.0); }
    // approx N(0,1): sum of 12 uniforms - 6, built from integer pieces
    double normal() {
        int64_t acc = 0;
        for (int i = 0; i < 6; ++i) { uint64_t r = next(); acc += (int64_t)(r & 0xffffff) + (int64_t)((r >> 24) & 0xffffff); }
        return (double)acc * (1.0 / 16777216.0) - 6.0;
    }
};
const char kBase[5] = {'A', 'C', 'G', 'T', 'N'};
inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
}  // namespace

void synth_reference(const SynthRefConfig &cfg, SynthRef &out) {
    out = SynthRef();
    out.cfg = cfg;
    Rng rng(cfg.seed);
    for (int c = 1; c <= 22; ++c) out.chrom_names.push_back(std::to_string(c));
    out.chrom_names.push_back("X");
    out.chrom_names.push_back("Y");
    const int n_auto = cfg.n_long + cfg.n_short;
    std::vector<int> per_chrom(24, 0);
    for (int j = 0; j < n_auto; ++j) ++per_chrom[j % 22];
    per_chrom[22] = cfg.n_x; per_chrom[23] = cfg.n_y;
    const int lead = cfg.flank_long + 1000;
    out.chrom_seq.resize(24);
    for (int c = 0; c < 24; ++c) {
        size_t len = (size_t)lead * 2 + (size_t)cfg.spacing * (size_t)std::max(per_chrom[c], 1);
        std::string &s = out.chrom_seq[c];
        s.resize(len);
        for (size_t i = 0; i < len; i += 32) {
            uint64_t r = rng.next();
            for (size_t k = 0; k < 32 && i + k < len; ++k, r >>= 2) s[i + k] = kBase[r & 3];
        }
    }
    // Planted repeats: a random genome has none, so the paths that only repeats reach (several occurrences per SA interval,
    // bwa_aln2seq_core's random pick, mapQ 0, pairing over many positions, XA lists) would never run.  Each copy takes the
    // 200 bases that start 20 bases right of one marker and writes them at the same offset of another marker; its own
    // generator keeps the main stream, and with it every existing fixture, unchanged.
    if (cfg.n_dup > 0) {
        Rng dup(cfg.seed ^ 0xD0B1E5EEDull);
        std::vector<std::pair<int, int>> sites;                  // (chromosome, 0-based marker position)
        for (int c = 0; c < 24; ++c) for (int m = 0; m < per_chrom[c]; ++m) sites.emplace_back(c, lead + cfg.spacing * m);
        for (int d = 0; d < cfg.n_dup && sites.size() >= 2; ++d) {
            const size_t a = dup.below((uint32_t)sites.size());
            size_t b = dup.below((uint32_t)sites.size() - 1);
            if (b >= a) ++b;
            const std::string seg = out.chrom_seq[sites[a].first].substr((size_t)sites[a].second + 20, 200);
            out.chrom_seq[sites[b].first].replace((size_t)sites[b].second + 20, 200, seg);
        }
    }
    // markers in genome (VCF) order; the first n_long autosomal records become long (RefBuilder::IsMaxNumMarker)
    int n_long_seen = 0;
    for (int c = 0; c < 24; ++c) {
        for (int m = 0; m < per_chrom[c]; ++m) {
            SynthMarker mk;
            mk.chrom = out.chrom_names[c];
            mk.pos = lead + cfg.spacing * m + 1;
            mk.is_long = (c < 22) && (n_long_seen < cfg.n_long);
            if (mk.is_long) ++n_long_seen;
            mk.flank = mk.is_long ? cfg.flank_long : cfg.flank_short;
            std::string &s = out.chrom_seq[c];
            mk.ref = s[mk.pos - 1];
            mk.alt = kBase[(std::find(kBase, kBase + 4, mk.ref) - kBase + 1 + rng.below(3)) & 3];
            mk.af = 0.05 + 0.9 * (double)rng.below(1000) / 1000.0;
            mk.flank_seq = s.substr((size_t)(mk.pos - 1 - mk.flank), (size_t)(2 * mk.flank + 1));
            mk.gc.resize((size_t)(2 * mk.flank + 1));
            // CalculateGC (src/RefBuilder.cpp:38-54): window = 1-based [i-50, i+49]
            int64_t w0 = (int64_t)mk.pos - mk.flank - 50 - 1;
            int run = 0;
            for (int k = 0; k < 100; ++k) { char ch = s[(size_t)(w0 + k)]; run += (ch == 'G' || ch == 'C'); }
            for (int t = 0; t < 2 * mk.flank + 1; ++t) {
                mk.gc[t] = (uint8_t)run;
                char o = s[(size_t)(w0 + t)], n = s[(size_t)(w0 + t + 100)];
                run += (n == 'G' || n == 'C') - (o == 'G' || o == 'C');
            }
            int off;
            do { off = (int)rng.below((uint32_t)(2 * mk.flank + 1)) - mk.flank; } while (off == 0);
            mk.extra_snp_pos = mk.pos + off;
            out.markers.push_back(std::move(mk));
        }
    }
    // flank FASTA order = std::map<string, map<int,...>> iteration (RefBuilder::PrepareRefSeq)
    out.flank_order.resize(out.markers.size());
    for (size_t i = 0; i < out.markers.size(); ++i) out.flank_order[i] = (int)i;
    std::stable_sort(out.flank_order.begin(), out.flank_order.end(), [&](int a, int b) {
        const SynthMarker &x = out.markers[a], &y = out.markers[b];
        if (x.chrom != y.chrom) return x.chrom < y.chrom;
        return x.pos < y.pos;
    });
}

static std::string marker_contig_name(const SynthMarker &m) {
    char buf[128];
    snprintf(buf, sizeof buf, "%s:%d@%c/%c%s", m.chrom.c_str(), m.pos, m.ref, m.alt, m.is_long ? "|L" : "");
    return buf;
}

std::vector<FlankSeq> synth_flanks(const SynthRef &ref) {
    std::vector<FlankSeq> v;
    v.reserve(ref.markers.size());
    for (int i : ref.flank_order) {
        const SynthMarker &m = ref.markers[i];
        std::string s = m.flank_seq;
        s[(size_t)m.flank] = m.ref;
        v.push_back(FlankSeq{marker_contig_name(m), s});
    }
    return v;
}

static const char *kVcfHeader =
    "##fileformat=VCFv4.1\n"
    "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency\">\n"
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n";

static void vcf_line(FILE *fp, const SynthMarker &m, int idx, bool tag_long) {
    fprintf(fp, "%s\t%d\trs%d%s\t%c\t%c\t.\tPASS\tAF=%.3f\n", m.chrom.c_str(), m.pos, idx + 1,
            (tag_long && m.is_long) ? "|L" : "", m.ref, m.alt, m.af);
}

static void write_dbsnp(FILE *fp, const SynthRef &ref) {
    fputs(kVcfHeader, fp);
    struct Rec { int chrom_i; int pos; char r, a; int id; };
    std::vector<Rec> recs;
    std::map<std::string, int> ci;
    for (size_t c = 0; c < ref.chrom_names.size(); ++c) ci[ref.chrom_names[c]] = (int)c;
    int id = 0;
    for (const SynthMarker &m : ref.markers) {
        int c = ci[m.chrom];
        recs.push_back(Rec{c, m.pos, m.ref, m.alt, ++id});
        char r = ref.chrom_seq[c][(size_t)(m.extra_snp_pos - 1)];
        recs.push_back(Rec{c, m.extra_snp_pos, r, comp(r), ++id});
    }
    std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) { return a.chrom_i != b.chrom_i ? a.chrom_i < b.chrom_i : a.pos < b.pos; });
    for (const Rec &r : recs)
        fprintf(fp, "%s\t%d\tdb%d\t%c\t%c\t.\tPASS\t.\n", ref.chrom_names[r.chrom_i].c_str(), r.pos, r.id, r.r, r.a);
}

bool synth_write_reference_inputs(const SynthRef &ref, const std::string &dir, std::string &err) {
    std::string fa = dir + "/genome.fa";
    FILE *fp = fopen(fa.c_str(), "w"), *fai = fopen((fa + ".fai").c_str(), "w"), *amb = fopen((fa + ".amb").c_str(), "w");
    if (!fp || !fai || !amb) { err = "cannot write under " + dir; return false; }
    long long off = 0, total = 0;
    for (size_t c = 0; c < ref.chrom_names.size(); ++c) {
        const std::string &s = ref.chrom_seq[c];
        off += fprintf(fp, ">%s\n", ref.chrom_names[c].c_str());
        fprintf(fai, "%s\t%zu\t%lld\t60\t61\n", ref.chrom_names[c].c_str(), s.size(), off);
        for (size_t i = 0; i < s.size(); i += 60) {
            size_t n = std::min<size_t>(60, s.size() - i);
            fwrite(s.data() + i, 1, n, fp); fputc('\n', fp);
            off += (long long)n + 1;
        }
        total += (long long)s.size();
    }
    fprintf(amb, "%lld %zu 0\n", total, ref.chrom_names.size());
    fclose(fp); fclose(fai); fclose(amb);
    fp = fopen((dir + "/markers.vcf").c_str(), "w");
    if (!fp) { err = "cannot write markers.vcf"; return false; }
    fputs(kVcfHeader, fp);
    for (size_t i = 0; i < ref.markers.size(); ++i) vcf_line(fp, ref.markers[i], (int)i, false);
    fclose(fp);
    fp = fopen((dir + "/dbsnp.vcf").c_str(), "w");
    if (!fp) { err = "cannot write dbsnp.vcf"; return false; }
    write_dbsnp(fp, ref);
    fclose(fp);
    return true;
}

bool synth_write_index_side_files(const SynthRef &ref, const std::string &genome_path, const std::string &dbsnp_path,
                                  const std::string &prefix, std::string &err) {
    FILE *fa = fopen(prefix.c_str(), "w"), *gc = fopen((prefix + ".gc").c_str(), "wb"),
         *vcf = fopen((prefix + ".SelectedSite.vcf").c_str(), "w"), *bed = fopen((prefix + ".bed").c_str(), "w"),
         *db = fopen((prefix + ".dbSNP.subset.vcf").c_str(), "w"), *par = fopen((prefix + ".param").c_str(), "w");
    if (!fa || !gc || !vcf || !bed || !db || !par) { err = "cannot write index side files for " + prefix; return false; }
    fputs(kVcfHeader, vcf);
    for (int i : ref.flank_order) {
        const SynthMarker &m = ref.markers[i];
        std::string s = m.flank_seq;
        s[(size_t)m.flank] = m.ref;
        fprintf(fa, ">%s\n%s\n", marker_contig_name(m).c_str(), s.c_str());
        uint32_t len = (uint32_t)m.gc.size();
        fwrite(&len, 4, 1, gc);
        fwrite(m.gc.data(), 1, len, gc);
        vcf_line(vcf, m, i, true);
        fprintf(bed, "%s\t%d\t%d\n", m.chrom.c_str(), m.pos - m.flank, m.pos + m.flank);
    }
    write_dbsnp(db, ref);
    fprintf(par, "REFERENCE_PATH\t%s\nTARGET_REGION_PATH\tEmpty\nDBSNP_VCF_PATH\t%s\nNUM_VAR_LONG\t%d\nNUM_VAR_SHORT\t%d\n"
                 "SHORT_FLANK_LENGTH\t%d\nLONG_FLANK_LENGTH\t%d\n",
            genome_path.c_str(), dbsnp_path.c_str(), ref.cfg.n_long, ref.cfg.n_short, ref.cfg.flank_short, ref.cfg.flank_long);
    fclose(fa); fclose(gc); fclose(vcf); fclose(bed); fclose(db); fclose(par);
    return true;
}

// ------------------------------------------------------------------------
static void make_read(Rng &rng, const SynthReadConfig &cfg, const std::string &frag, bool rc, uint8_t *bases, uint8_t *quals) {
    // walk the fragment from the 5' end of this read, injecting errors
    const int L = cfg.read_len, F = (int)frag.size();
    int fp = 0, o = 0;
    auto tmpl = [&](int k) -> char { return rc ? comp(frag[(size_t)(F - 1 - k)]) : frag[(size_t)k]; };
    while (o < L) {
        if (fp >= F) { bases[o++] = (uint8_t)kBase[rng.below(4)]; continue; }
        double u = rng.unif();
        if (u < cfg.del_rate) { fp += 1 + (cfg.max_indel_len > 1 ? (int)rng.below((uint32_t)cfg.max_indel_len) : 0); continue; }
        if (u < cfg.del_rate + cfg.ins_rate) {
            int n = 1 + (cfg.max_indel_len > 1 ? (int)rng.below((uint32_t)cfg.max_indel_len) : 0);
            for (int k = 0; k < n && o < L; ++k) bases[o++] = (uint8_t)kBase[rng.below(4)];
            continue;
        }
        char c = tmpl(fp++);
        if (u < cfg.del_rate + cfg.ins_rate + cfg.sub_rate) {
            int b = (int)(std::find(kBase, kBase + 4, c) - kBase);
            c = kBase[(b + 1 + (int)rng.below(3)) & 3];
        }
        bases[o++] = (uint8_t)c;
    }
    int tail_from = L;
    if (rng.unif() < cfg.bad_tail_rate) tail_from = L / 3 + (int)rng.below((uint32_t)(L - L / 3));
    for (int k = 0; k < L; ++k) {
        if (rng.unif() < cfg.n_rate) bases[k] = 'N';
        double q = (k >= tail_from ? 9.0 : 36.0 - 0.08 * k) + 4.0 * rng.normal();
        int qi = (int)std::floor(q + 0.5);
        qi = std::max(2, std::min(41, qi));
        quals[k] = (uint8_t)(33 + qi);
    }
}

static void one_pair(const SynthRef &ref, const SynthReadConfig &cfg, int64_t pair, uint8_t *b1, uint8_t *q1, uint8_t *b2, uint8_t *q2) {
    Rng rng(cfg.seed * 0x9E3779B97F4A7C15ull + (uint64_t)pair * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull);
    const int L = cfg.read_len;
    std::string frag;
    if (rng.unif() < cfg.f_on) {
        const SynthMarker &m = ref.markers[rng.below((uint32_t)ref.markers.size())];
        int span = 2 * m.flank + 1;
        int isz = (int)std::floor(cfg.isize_mean + cfg.isize_sd * rng.normal() + 0.5);
        isz = std::max(L + 10, std::min(span, isz));
        int start = (int)rng.below((uint32_t)(span - isz + 1));
        frag = m.flank_seq.substr((size_t)start, (size_t)isz);
        int mpos = m.flank - start;
        if (mpos >= 0 && mpos < isz) frag[(size_t)mpos] = (rng.unif() < m.af) ? m.alt : m.ref;
    } else {
        int isz = (int)std::floor(cfg.isize_mean + cfg.isize_sd * rng.normal() + 0.5);
        isz = std::max(L + 10, isz);
        frag.resize((size_t)isz);
        for (char &c : frag) c = kBase[rng.below(4)];
    }
    bool swap = rng.below(2) != 0;
    // R1 reads the fragment forward, R2 its reverse complement; `swap` flips which end is which
    make_read(rng, cfg, frag, swap, b1, q1);
    make_read(rng, cfg, frag, !swap, b2, q2);
}

void synth_reads(const SynthRef &ref, const SynthReadConfig &cfg, int64_t first_pair, int64_t n_pairs,
                 uint8_t *bases1, uint8_t *quals1, uint8_t *bases2, uint8_t *quals2, int n_threads) {
    const int L = cfg.read_len;
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_pairs / 1024));
    auto work = [&](int t) {
        int64_t lo = n_pairs * t / n_threads, hi = n_pairs * (t + 1) / n_threads;
        for (int64_t i = lo; i < hi; ++i)
            one_pair(ref, cfg, first_pair + i, bases1 + i * L, quals1 + i * L, bases2 + i * L, quals2 + i * L);
    };
    if (n_threads == 1) { work(0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
}

}  // namespace fqb
