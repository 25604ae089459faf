// See fq_stats_host.h.  Output formatting goes through std::ostream exactly like the
// reference's writers, so doubles print with the same default 6-significant-digit rule.
#include "fq_stats_host.h"
#include "fq_common.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <unordered_map>

namespace fqb {

static std::string norm_chrom(std::string c) {
    std::transform(c.begin(), c.end(), c.begin(), ::toupper);
    if (c.find("CHR") != std::string::npos) c = c.substr(3);
    return c;
}

static bool read_vcf(const std::string &path, std::vector<MarkerRec> &out, std::string &err) {
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> f;
        size_t b = 0;
        while (true) {
            size_t e = line.find('\t', b);
            f.push_back(line.substr(b, e == std::string::npos ? std::string::npos : e - b));
            if (e == std::string::npos) break;
            b = e + 1;
        }
        if (f.size() < 5) continue;
        MarkerRec m;
        m.chrom_raw = f[0]; m.chrom = norm_chrom(f[0]); m.pos = atoi(f[1].c_str());
        m.id = f[2]; m.ref = f[3]; m.alt = f[4];
        m.qual = f.size() > 5 ? f[5] : "."; m.filter = f.size() > 6 ? f[6] : ".";
        if (f.size() > 7) {
            std::string info = ";" + f[7];
            size_t p = info.find(";AF=");
            if (p != std::string::npos) { size_t e = info.find(';', p + 1); m.af = info.substr(p + 4, e == std::string::npos ? std::string::npos : e - p - 4); m.has_af = true; }
        }
        out.push_back(m);
    }
    return true;
}

// RegionList::AddRegion + Collapse (src/RegionList.cpp:65-118): membership = union of the kept intervals
struct Regions {
    std::map<std::string, std::map<int, int>> r;
    void add(const std::string &chr, int s, int e) { r[norm_chrom(chr)][s] = e; }
    void collapse() {
        std::map<std::string, std::map<int, int>> t;
        for (auto kv : r) {
            auto holder = kv.second.begin();
            for (auto it = kv.second.begin(); it != kv.second.end(); ++it) {
                int b1 = holder->first, e1 = holder->second, b2 = it->first, e2 = it->second;
                if (e1 >= e2) continue;
                else if (e1 < b2) { t[kv.first][b1] = e1; holder = it; }
                else { t[kv.first][b1] = e2; holder->second = e2; }
            }
            t[kv.first][holder->first] = holder->second;
        }
        r = t;
    }
    uint64_t size() const {                 // RegionList::Size(): set by Collapse
        uint64_t len = 0;
        for (auto &kv : r) for (auto &iv : kv.second) len += (uint64_t)(iv.second - iv.first) + 1;
        return len;
    }
    // RegionList::ReadRegionList(path, autoCollapse = true) (src/RegionList.cpp:15-45)
    bool read_bed(const std::string &path) {
        std::ifstream fin(path);
        if (!fin.is_open()) return false;
        std::string line;
        while (std::getline(fin, line)) {
            std::string chr;
            int start = 0, end = 0;
            std::stringstream ss(line);
            ss >> chr >> start >> end;
            chr = norm_chrom(chr);
            auto &m = r[chr];
            auto it = m.find(start);
            if (it != m.end()) { if (it->second < end) it->second = end; } else m[start] = end;
        }
        collapse();
        return true;
    }
    // RegionList::Join(b, isUnion = false) (src/RegionList.cpp:120-173)
    void inner_join(const Regions &b) {
        collapse();
        Regions t;
        for (auto &kv : b.r) {
            auto mine = r.find(kv.first);
            if (mine == r.end()) continue;
            auto a = mine->second.begin();
            auto it = kv.second.begin();
            while (a != mine->second.end() && it != kv.second.end()) {
                const int b1 = a->first, e1 = a->second, b2 = it->first, e2 = it->second;
                if (b1 <= b2) {
                    if (e1 > e2) { t.r[kv.first][b2] = e2; ++it; }
                    else if (e1 > b2) { t.r[kv.first][b2] = e1; ++a; }
                    else ++a;
                } else {
                    if (e1 <= e2) { t.r[kv.first][b1] = e1; ++a; }
                    else if (e1 > b2 && b1 < e2) { t.r[kv.first][b1] = e2; ++it; }
                    else ++it;
                }
            }
        }
        r = t.r;
        collapse();
    }
    bool has(const std::string &chr, int pos) const {
        auto c = r.find(chr);
        if (c == r.end()) return false;
        auto it = c->second.lower_bound(pos);
        if (it != c->second.end() && it->first <= pos && it->second >= pos) return true;
        if (it != c->second.begin()) { --it; if (it->first <= pos && it->second >= pos) return true; }
        return false;
    }
};

bool build_stats_tables(const HostIndex &idx, const std::string &prefix, const fqb_gap_opt_t &g, const std::string &target_bed, StatsTables &T,
                        std::string &err) {
    T = StatsTables();
    if (!read_vcf(prefix + ".SelectedSite.vcf", T.markers, err)) return false;
    std::ifstream fgc(prefix + ".gc", std::ios::binary);
    if (!fgc) { err = "cannot open " + prefix + ".gc"; return false; }
    T.chopped_read_len = (int)std::floor(g.read_len * 0.65f + 0.5);        // FLANK_EDGE, src/StatCollector.cpp:28,1753
    std::map<std::string, std::map<int, unsigned>> vcf_table;              // VcfTable
    std::unordered_map<std::string, std::unordered_map<int, unsigned>> gc, dbsnp;
    Regions flank;
    for (size_t i = 0; i < T.markers.size(); ++i) {
        const MarkerRec &m = T.markers[i];
        vcf_table[m.chrom][m.pos] = (unsigned)i;
        uint32_t len = 0;
        fgc.read(reinterpret_cast<char *>(&len), 4);
        std::vector<unsigned char> buf(len);
        if (len) fgc.read(reinterpret_cast<char *>(buf.data()), len);
        int tmp_pos = m.pos - (int)(len - 1) / 2;
        for (uint32_t k = 0; k < len; ++k) gc[m.chrom][tmp_pos + (int)k] = buf[k];
        int fl;
        if (m.chrom == "X" || m.chrom == "Y") { ++T.n_xy; fl = g.flank_len; }
        else if (!m.id.empty() && m.id.back() == 'L') { ++T.n_long; fl = g.flank_long_len; }
        else { ++T.n_short; fl = g.flank_len; }
        flank.add(m.chrom, m.pos - fl + T.chopped_read_len, m.pos + fl - T.chopped_read_len);
    }
    flank.collapse();
    if (!target_bed.empty()) {              // StatCollector::SetTargetRegion (src/StatCollector.cpp:2284-2288)
        Regions target;
        if (!target.read_bed(target_bed)) { err = "Region list bed file:" + target_bed + " open failed!"; return false; }
        flank.inner_join(target);
        T.has_target = true; T.flank_region_size = flank.size(); T.target_region_size = target.size();
    }
    {
        std::vector<MarkerRec> db;
        if (!read_vcf(prefix + ".dbSNP.subset.vcf", db, err)) return false;
        for (const MarkerRec &m : db) dbsnp[m.chrom][m.pos] = 1;
    }
    for (auto &c : vcf_table) for (auto &p : c.second) T.marker_out_order.push_back((int)p.second);
    // contigs: "chr:pos@R/A[|L]" -> genome coordinates (src/StatCollector.cpp:453-481)
    const size_t l_pac = (size_t)idx.l_pac;
    T.site.assign(l_pac, kSiteNone);
    T.marker_at.assign(l_pac, -1);
    std::map<std::pair<std::string, int>, uint32_t> site_ids;
    for (const Contig &c : idx.contigs) {
        ContigDev d;
        memset(&d, 0, sizeof d);
        d.offset = c.offset; d.len = c.len;
        size_t colon = c.name.find(':');
        if (colon == std::string::npos) { err = "contig name without ':' (external alignments are not supported): " + c.name; return false; }
        std::string chrom = norm_chrom(c.name.substr(0, colon));
        int ref_coord = (int)strtol(c.name.c_str() + colon + 1, nullptr, 10);
        int fl = c.name.back() == 'L' ? g.flank_long_len : g.flank_len;
        d.gstart = ref_coord - fl;
        d.is_xy = (c.name.find('X') != std::string::npos || c.name.find('Y') != std::string::npos) ? 1 : 0;
        T.contigs.push_back(d);
        T.contig_names.push_back(c.name);
        auto vt = vcf_table.find(chrom);
        auto dt = dbsnp.find(chrom);
        auto gt = gc.find(chrom);
        for (int k = 0; k < c.len; ++k) {
            const int gpos = d.gstart + k;
            const size_t x = (size_t)c.offset + (size_t)k;
            uint32_t v = kSiteNone;
            if (flank.has(chrom, gpos)) {
                auto key = std::make_pair(chrom, gpos);
                auto it = site_ids.find(key);
                if (it == site_ids.end()) {
                    it = site_ids.emplace(key, T.n_sites++).first;
                    unsigned char gcv = 0;
                    if (gt != gc.end()) { auto gi = gt->second.find(gpos); if (gi != gt->second.end()) gcv = (unsigned char)gi->second; }
                    T.site_gc.push_back(gcv);
                }
                v = it->second;
            }
            if (dt != dbsnp.end() && dt->second.count(gpos)) v |= kSiteDbsnp;
            if (vt != vcf_table.end()) {
                auto mi = vt->second.find(gpos);
                if (mi != vt->second.end()) { v |= kSiteMarker; T.marker_at[x] = (int32_t)mi->second; }
            }
            T.site[x] = v;
        }
    }
    // BwtIndexer::LoadContigSize (src/BwtIndexer.cpp:764-802): genome size from <ref>.fai, "N size" from <ref>.amb
    {
        std::ifstream par(prefix + ".param");
        std::string k, ref_path;
        par >> k >> ref_path;
        std::ifstream fai(ref_path + ".fai");
        std::string line;
        while (std::getline(fai, line)) { std::stringstream ss(line); std::string chr, len; ss >> chr >> len; T.ref_genome_size += (uint64_t)atoi(len.c_str()); T.genome_contigs.emplace_back(chr, atoi(len.c_str())); }
        std::ifstream amb(ref_path + ".amb");
        while (std::getline(amb, line)) { std::stringstream ss(line); std::string off, nlen; ss >> off >> nlen; T.ref_N_size += (uint64_t)atoi(nlen.c_str()); }
    }
    return true;
}

// decimal append without iostreams: the table has ~1 line per pair, so this is the host hot spot of the statistics stage
static inline void put_int(std::string &o, long long v) {
    char buf[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) o.push_back('-');
    while (n) o.push_back(buf[--n]);
}
static inline void put_cigar(std::string &o, const fqb_read_t &p) {      // Cigar2String (src/StatCollector.cpp:56-71)
    if (!p.has_cigar) { put_int(o, p.len); o.push_back('M'); return; }
    for (int k = 0; k < p.n_cigar; ++k) { put_int(o, p.cigar[k] & 0x3fff); o.push_back("MIDS"[p.cigar[k] >> 14]); }
}

// appends the line (nothing when the pair prints none)
void append_isize_line(const StatsTables &T, const PairStat &ps, const fqb_read_t &p, const fqb_read_t &q, const char *name, std::string &o) {
    static const char *kStatus[] = {"", "PropPair", "PartialPair", "NotPair", "LowQual", "FwdOnly", "RevOnly"};
    if (ps.line_kind == 0) return;
    o += name; o.push_back('\t');
    put_int(o, ps.max_insert); o.push_back('\t'); put_int(o, ps.max_insert2); o.push_back('\t'); put_int(o, ps.actual_insert); o.push_back('\t');
    const fqb_read_t *rd[2] = {&p, &q};
    for (int e = 0; e < 2; ++e) {
        if (ps.line_kind & (1 << e)) {
            const ContigDev &c = T.contigs[ps.seqid[e]];
            o += T.contig_names[ps.seqid[e]]; o.push_back('\t');
            put_int(o, (long long)rd[e]->pos - c.offset + 1); o.push_back('\t');
            put_int(o, ps.flag[e]); o.push_back('\t');
            put_int(o, rd[e]->len); o.push_back('\t');
            put_cigar(o, *rd[e]); o.push_back('\t');
        } else {
            o += "*\t*\t"; put_int(o, ps.flag[e]); o += "\t0\t*\t";
        }
    }
    o += kStatus[ps.status]; o.push_back('\n');
}
void format_isize_line(const StatsTables &T, const PairStat &ps, const fqb_read_t &p, const fqb_read_t &q, const char *name, std::string &out) {
    out.clear();
    append_isize_line(T, ps, p, q, name, out);
}

// ---- InsertSizeEstimator (src/InsertSizeEstimator.cpp:43-173) -------------------------------
static std::vector<double> adjusted_isize(const std::string &table, const std::string &orientation) {
    const int LIM = 4096;
    std::vector<double> mis(LIM, 1e-6), obs(LIM, 1e-6);     // initEp (src/InsertSizeEstimator.h:60,80-81)
    int total = 0;
    std::ifstream fin(table);
    std::string line;
    while (std::getline(fin, line)) {
        std::vector<std::string> f;
        size_t b = 0;
        while (true) { size_t e = line.find('\t', b); f.push_back(line.substr(b, e == std::string::npos ? std::string::npos : e - b)); if (e == std::string::npos) break; b = e + 1; }
        if (f.size() < 15) continue;
        int Max = atoi(f[1].c_str()), Max2 = atoi(f[2].c_str()), Obs = atoi(f[3].c_str()), Flag1 = atoi(f[6].c_str()), Flag2 = atoi(f[11].c_str());
        const std::string &c1 = f[8], &c2 = f[13], &st = f[14];
        if (Max >= LIM || Max == -1) Max = LIM - 1;
        if (Max2 >= LIM || Max2 == -1) Max2 = LIM - 1;
        if (Obs >= LIM || Obs == -1) Obs = LIM - 1;
        if (st == "Abnormal" || st == "LowQual" || st == "NotPair" || st == orientation) continue;
        else if (st == "FwdOnly") mis[Max] += 1.;
        else if (st == "RevOnly") mis[Max2] += 1.;
        else if (st == "PropPair") obs[Obs] += 1.;
        else if (st == "PartialPair") {
            if (c1.find('S') == std::string::npos && c2.find('S') != std::string::npos) { if (Flag1 & 16) mis[Max2] += 1.; else mis[Max] += 1.; }
            else if (c1.find('S') != std::string::npos && c2.find('S') == std::string::npos) { if (Flag2 & 16) mis[Max2] += 1.; else mis[Max] += 1.; }
            else continue;
        } else exit(EXIT_FAILURE);
        ++total;
    }
    std::vector<double> F(2000, 0.), f(2000, 0.), G(2000, 0.), gg(2000, 0.);
    for (int k = 0; k < 2000; ++k) {
        double m = mis[k], n = obs[k];
        if (k != 0) { f[k] = n / (1 - G[k - 1]) * 1 / double(total); F[k] = F[k - 1] + f[k]; }
        else { f[k] = n / double(total); F[k] = f[k]; }
        if (k != 0) { gg[k] = m / (1 - F[k]) * 1 / double(total); G[k] = G[k - 1] + gg[k]; }
        else { gg[k] = m / double(total); G[k] = gg[k]; }
    }
    return f;
}

// GetInsertSizeDist's first half (src/StatCollector.cpp:1969-1985): both censoring directions, summed
bool write_adjusted_isize(const std::string &table, const std::string &out_path) {
    std::vector<double> f1 = adjusted_isize(table, "FwdOnly");
    std::vector<double> f2 = adjusted_isize(table, "RevOnly");
    std::ofstream fa(out_path);
    for (size_t i = 0; i < f1.size(); ++i) { f1[i] = (f1[i] + f2[i]); fa << i << "\t" << f1[i] << std::endl; }
    return (bool)fa;
}

// the reference calls these unqualified under `using namespace std`, i.e. with the float overloads where the argument is float
#define REV_PHRED(x) std::pow(10.0, (x / (-10.0)))
#define PHRED(x) (-10) * std::log10(x)

// CalLikelihood (src/StatCollector.cpp:2069-2096): same float/double mix as the reference
static std::vector<float> cal_likelihood(const std::string &seq, const std::string &qual, char maj, char min) {
    float GL0(0), GL1(0), GL2(0);
    for (uint32_t i = 0; i != seq.size(); ++i) {
        float seq_error = REV_PHRED(qual[i]);
        if (seq[i] == maj) { GL0 += std::log10(1 - seq_error); GL1 += std::log10(0.5 - seq_error / 3); GL2 += std::log10(seq_error / 3); }
        else if (seq[i] == min) { GL0 += std::log10(seq_error / 3); GL1 += std::log10(0.5 - seq_error / 3); GL2 += std::log10(1 - seq_error); }
        else { GL0 += std::log10(2 * seq_error / 3); GL1 += std::log10(2 * seq_error / 3); GL2 += std::log10(2 * seq_error / 3); }
    }
    std::vector<float> t(3, 0);
    t[0] = std::floor(GL0 * (-10) + 0.5); t[1] = std::floor(GL1 * (-10) + 0.5); t[2] = std::floor(GL2 * (-10) + 0.5);
    return t;
}

bool write_summary_files(const StatsTables &T, StatsTotals &S, const fqb_gap_opt_t &g, const std::string &prefix, std::string &err) {
    using std::endl;
    // ---- GetDepthDist (1858-1914)
    std::vector<size_t> DepthDist(1024, 0), GCDist(256, 0), PosNum(101, 0);
    uint64_t NumBaseMapped = 0, Cov = 0, Cov2 = 0, Cov5 = 0, Cov10 = 0;
    for (uint32_t s = 0; s < T.n_sites; ++s) {
        uint32_t d = S.depth[s];
        if (d == 0) continue;                         // PositionTable only holds sites that were touched
        NumBaseMapped += d;
        DepthDist[d > 1023 ? 1023 : d]++;
        GCDist[T.site_gc[s]] += d;
        if (T.site_gc[s] < 101) PosNum[T.site_gc[s]]++;
    }
    for (size_t i = 1; i != DepthDist.size(); ++i) {
        Cov += DepthDist[i];
        if (i >= 2) Cov2 += DepthDist[i];
        if (i >= 5) Cov5 += DepthDist[i];
        if (i >= 10) Cov10 += DepthDist[i];
    }
    const int ch = T.chopped_read_len;
    const uint64_t total_region_size = T.has_target ? T.flank_region_size :
                                       (uint64_t)(((g.flank_len - ch) * 2 + 1)) * T.n_short + (uint64_t)(((g.flank_long_len - ch) * 2 + 1)) * T.n_long +
                                       (uint64_t)(((g.flank_len - ch) * 2 + 1)) * T.n_xy;
    {
        std::ofstream f(prefix + ".DepthDist");
        if (!f) { err = "cannot write " + prefix + ".DepthDist"; return false; }
        DepthDist[0] = total_region_size - Cov;
        for (uint32_t i = 0; i != DepthDist.size(); ++i) f << i << "\t" << DepthDist[i] << endl;
    }
    {   // GetGCDist (1916-1932)
        std::ofstream f(prefix + ".GCDist");
        double MeanDepth = NumBaseMapped / (double)Cov;
        for (uint32_t i = 0; i != 101; ++i) {
            f << i << "\t" << GCDist[i] << "\t" << PosNum[i] << "\t";
            if (PosNum[i] == 0) f << 0; else f << (double(GCDist[i]) / PosNum[i]) / MeanDepth;
            f << endl;
        }
    }
    const unsigned long long *Emp = S.emp.data(), *misEmp = Emp + 256, *EmpCyc = Emp + 512, *misCyc = Emp + 768;
    {   // GetEmpRepDist (1934-1948)
        std::ofstream f(prefix + ".EmpRepDist");
        for (uint32_t i = 0; i != 256; ++i) {
            f << i << "\t" << (size_t)misEmp[i] << "\t" << (size_t)Emp[i] << "\t";
            if (Emp[i] == 0) f << 0; else f << PHRED((double)(misEmp[i] + 1) / (Emp[i] + 2));
            f << endl;
        }
    }
    {   // GetEmpCycleDist (1950-1967); CycleDist is never filled by the align stage
        std::ofstream f(prefix + ".EmpCycleDist");
        double prevQual = 0;
        for (uint32_t i = 0; i != 256; ++i) {
            f << i + 1 << "\t" << (size_t)misCyc[i] << "\t" << (size_t)EmpCyc[i] << "\t";
            if (misCyc[i] == 0) f << prevQual; else f << PHRED((double)(misCyc[i] + 1e-6) / (EmpCyc[i] + 1e-6));
            f << "\t" << (size_t)0 << endl;
            if (misCyc[i] != 0) prevQual = PHRED((double)(misCyc[i] + 1e-6) / (EmpCyc[i] + 1e-6));
        }
    }
    {   // GetInsertSizeDist (1969-1997)
        write_adjusted_isize(prefix + ".InsertSizeTable", prefix + ".AdjustedInsertSizeDist");
        std::ofstream fr(prefix + ".RawInsertSizeDist");
        for (uint32_t i = 0; i != 4096; ++i) fr << i << "\t" << (size_t)S.isize_dist[i] << endl;
    }
    {   // GetSexChromInfo (1999-2010): iteration order of a std::unordered_map<string, ...> filled in first-touch order
        std::vector<std::pair<uint32_t, int>> order;
        for (size_t c = 0; c < T.contigs.size(); ++c) if (S.contig_first[c] != 0xffffffffu) order.emplace_back(S.contig_first[c], (int)c);
        std::sort(order.begin(), order.end());
        std::unordered_map<std::string, int> table;
        for (auto &pr : order) table[T.contig_names[pr.second]] = pr.second;
        std::ofstream f(prefix + ".SexChromInfo");
        for (auto it = table.begin(); it != table.end(); ++it) {
            const uint32_t *c = &S.contig_ctr[4 * (size_t)it->second];
            f << it->first << "\t" << (int)c[0] << "\t" << (int)c[1] << "\t" << (int)c[2] << "\t" << (int)c[3] << endl;
        }
    }
    {   // GetPileup (2030-2066)
        std::ofstream f(prefix + ".Pileup");
        const int qualoffset = g.is_il13 ? 64 : 33;
        for (int mi : T.marker_out_order) {
            const PileupColumn &c = S.pileup[mi];
            if (c.seq.empty()) continue;
            f << T.markers[mi].chrom << "\t" << T.markers[mi].pos << "\t.\t" << c.strand.size() << "\t";
            for (uint32_t k = 0; k != c.strand.size(); ++k) f << (char)(c.strand[k] ? toupper(c.seq[k]) : tolower(c.seq[k]));
            f << "\t";
            for (uint32_t k = 0; k != c.qual.size(); ++k) f << char(c.qual[k] + qualoffset);
            f << "\t";
            for (uint32_t k = 0; k != c.maq.size(); ++k) f << c.maq[k];
            f << "\t";
            for (uint32_t k = 0; k != c.cycle.size(); ++k) { f << c.cycle[k]; if (k != c.cycle.size() - 1) f << ","; }
            f << endl;
        }
    }
    {   // SummaryOutput (2343-2483)
        std::ofstream fq(prefix + ".FASTQ.csv");
        fq << "FileIndex,PairEnd1,PairEnd2" << endl;
        for (size_t i = 0; i != S.files.size(); ++i) {
            auto strip = [](std::string &s) { size_t p = s.find_last_of("\\/"); if (p != std::string::npos) s.erase(0, p + 1); };
            strip(S.files[i].FileName1); strip(S.files[i].FileName2);
            fq << i + 1 << "," << S.files[i].FileName1 << "," << S.files[i].FileName2 << "\n";
        }
        fq.close();
        std::ofstream fc(prefix + ".Sequence.csv");
        long long total_base = 0, total_reads = 0, total_retained = 0, total_unmapped = 0, total_low = 0;
        fc << "FileIndex,NumOfBases,NumOfReads,NumOfUmappedReads,NumOfLowMAPQReads,NumOfQCPassReads,ReadLength" << endl;
        for (size_t i = 0; i != S.files.size(); ++i) {
            const FileCounters &F = S.files[i];
            fc << i + 1 << "," << F.NumBase << "," << F.NumRead << "," << F.BwaUnmapped << "," << F.TotalMAPQ << "," << F.TotalRetained << ",";
            fc << ((F.NumRead == 0) ? 0 : (F.NumBase / F.NumRead)) << endl;
            total_base += F.NumBase; total_reads += F.NumRead; total_retained += F.TotalRetained; total_unmapped += F.BwaUnmapped; total_low += F.TotalMAPQ;
        }
        double avgReadLen = std::floor(0.5 + ((total_reads == 0) ? 0 : ((double)total_base / total_reads)));
        fc << "Total," << total_base << "," << total_reads << "," << total_unmapped << "," << total_low << "," << total_retained << ",";
        fc << avgReadLen << endl;
        fc.close();
        std::ofstream f(prefix + ".Summary");
        f << "Statistics : " << "Value\n";
        auto report_genome_size = T.has_target ? T.target_region_size : (T.ref_genome_size - T.ref_N_size);
        double estimated_total_mapped_reads = (double)NumBaseMapped / avgReadLen * report_genome_size / total_region_size;
        f << "Estimated Read Mapping Rate : " << estimated_total_mapped_reads / total_reads << "\n";
        f << "Estimated Read PCR Duplication Rate : " << S.num_pcr_dup / ((double)S.num_pair_reads) << "[" << (uint64_t)S.num_pcr_dup << "/" << (double)S.num_pair_reads << "]\n";
        f << "Whole Genome Coverage : " << (double)total_base / T.ref_genome_size << "[" << total_base << "/" << T.ref_genome_size << "]\n";
        f << "Expected Read Depth : " << (double)total_base / report_genome_size << "[" << total_base << "/" << report_genome_size << "]\n";
        f << "Estimated Read Depth : ";
        if (Cov == 0) f << 0; else f << NumBaseMapped / (double)total_region_size;
        f << "[" << NumBaseMapped << "/" << total_region_size << "]\n";
        f << "Reduced Genome Size : " << total_region_size << endl;
        f << "Depth 1 or above position fraction : " << Cov / (double)total_region_size << endl;
        f << "Depth 2 or above position fraction : " << Cov2 / (double)total_region_size << endl;
        f << "Depth 5 or above position fraction : " << Cov5 / (double)total_region_size << endl;
        f << "Depth 10 or above position fraction : " << Cov10 / (double)total_region_size << endl;
        long long q20 = 0, q30 = 0;
        for (uint32_t s = 0; s < T.n_sites; ++s) { q20 += S.q20[s]; q30 += S.q30[s]; }
        f << "Q20 Base Fraction : " << (NumBaseMapped == 0 ? 0 : double(q20) / NumBaseMapped) << endl;
        f << "Q30 Base Fraction : " << (NumBaseMapped == 0 ? 0 : double(q30) / NumBaseMapped) << endl;
        f << "Estimated AvgDepth for Q20 bases : " << double(q20) / Cov << endl;
        f << "Estimated AvgDepth for Q30 bases : " << double(q30) / Cov << endl;
        auto mis = [&](size_t from) -> size_t {
            long long tmp = 0, total = 0;
            for (size_t i = from; i != 4096; ++i) total += (long long)S.isize_dist[i];
            for (size_t i = from; i != 4096; ++i) { tmp += (long long)S.isize_dist[i]; if (tmp > total / 2) return i; }
            return 0;
        };
        f << "Median Insert Size(>=500bp) : " << mis(500) << endl;
        f << "Median Insert Size(>=300bp) : " << mis(300) << endl;
    }
    {   // GetVCF (2185-2271)
        std::ofstream f(prefix + ".vcf");
        std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
        char buf[100] = {0};
        std::strftime(buf, sizeof(buf), "%Y%m%d", std::localtime(&now));
        f << "##fileformat=VCFv4.2\n" << "##fileDate=" << buf << "\n" << "##source=VerifyBamID2\n";
        f << "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency, for each ALT allele, in the same order as listed\">\n";
        f << "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
        f << "##FORMAT=<ID=GP,Number=1,Type=String,Description=\"Genotype\">\n";
        f << "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Normalized, Phred-scaled likelihoods for genotypes as defined in the VCF specification\">\n";
        f << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tIntendedSample\n";
        for (int mi : T.marker_out_order) {
            const MarkerRec &m = T.markers[mi];
            const PileupColumn &c = S.pileup[mi];
            if (!m.has_af) continue;
            float alleleFrq = atof(m.af.c_str());
            if (c.seq.empty()) continue;
            f << m.chrom_raw << "\t" << m.pos << "\t" << m.id << "\t" << m.ref << "\t" << m.alt << "\t" << m.qual << "\t" << m.filter << "\t";
            f << "AF=" << m.af << ";AC=" << c.seq.size() << "\t" << "GT:PL:GP\t";
            std::vector<float> pl = cal_likelihood(c.seq, c.qual, m.ref[0], m.alt[0]);
            float prior[3], post[3], sum;
            prior[0] = PHRED((1 - alleleFrq) * (1 - alleleFrq));
            prior[1] = PHRED(2 * alleleFrq * (1 - alleleFrq));
            prior[2] = PHRED(alleleFrq * alleleFrq);
            post[0] = prior[0] + pl[0]; post[1] = prior[1] + pl[1]; post[2] = prior[2] + pl[2];
            sum = PHRED(REV_PHRED(post[0]) + REV_PHRED(post[1]) + REV_PHRED(post[2]));
            post[0] = std::floor(post[0] - sum + 0.5); post[1] = std::floor(post[1] - sum + 0.5); post[2] = std::floor(post[2] - sum + 0.5);
            const char *gt = post[0] < post[1] ? (post[0] < post[2] ? "0/0:" : "1/1:") : (post[1] < post[2] ? "0/1:" : "1/1:");
            f << gt << pl[0] << "," << pl[1] << "," << pl[2] << ":" << post[0] << "," << post[1] << "," << post[2] << "\n";
        }
    }
    (void)err;
    return true;
}

}  // namespace fqb

// InsertSizeEstimator on a finished InsertSizeTable (host only; what fqb_stats_finish runs for <prefix>.AdjustedInsertSizeDist)
extern "C" int fqb_isize_adjusted_file(const char *table_path, const char *out_path) {
    if (!table_path || !out_path) { fqb::set_error("null argument"); return FQB_ERR_ARG; }
    { std::ifstream probe(table_path); if (!probe) { fqb::set_error(std::string("cannot read ") + table_path); return FQB_ERR_IO; } }
    if (!fqb::write_adjusted_isize(table_path, out_path)) { fqb::set_error(std::string("cannot write ") + out_path); return FQB_ERR_IO; }
    return FQB_OK;
}
