// TEST INFRASTRUCTURE ONLY.  Instantiates the per-lane device functions of
// fastquick_b200/csrc/fq_device_core.cuh on the host, one "lane" at a time, so
// the kernel LOGIC can be checked against the oracle where no GPU exists.  The
// product library never contains or calls this.
#define FQB_LANE_STATS 1
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include "../../fastquick_b200/csrc/fq_device_core.cuh"
#include "../../fastquick_b200/csrc/fq_device_pair.cuh"
#include "../../fastquick_b200/csrc/fq_device_dp.cuh"
#include <cmath>
#include "../../fastquick_b200/csrc/fq_hostmath.h"
#include "../../fastquick_b200/csrc/fq_index.h"
#include "../../fastquick_b200/csrc/fq_relayout.h"
#include "../../fastquick_b200/csrc/fq_stats_host.h"

using namespace fqb;

namespace fqb { void set_error(const std::string &) {} }     // the product's error slot lives in fq_capi_host.cpp, which the emulator does not link

static unsigned long long g_stats[40];
extern "C" void emul_stats(unsigned long long *o) { for (int i = 0; i < 40; ++i) { o[i] = g_stats[i]; g_stats[i] = 0; } }
// emul_search_var(1): the fast-pass lane is instantiated with pop staging (SearchLane kVar bit 0: the staging slot and its tag
// logic are the lane's own code, the asynchronous copy becomes an immediate one); emul_stage_stats: pops served from the slot,
// pops that went to the arena, staged entries that differed from the arena's when used (must be 0), since the last call
static int g_search_var = 0;
static unsigned long long g_stage[3];
extern "C" void emul_search_var(int v) { g_search_var = v; }
extern "C" void emul_stage_stats(unsigned long long *o) { for (int i = 0; i < 3; ++i) { o[i] = g_stage[i]; g_stage[i] = 0; } }
struct Emul {
    HostIndex idx;
    std::vector<Block32> blocks[2];
    DevBwt bwt[2];
};

extern "C" {

void *emul_open(const char *prefix, char *err, int errlen) {
    Emul *e = new Emul();
    std::string msg;
    if (!load_index(prefix, false, e->idx, msg)) { strncpy(err, msg.c_str(), errlen - 1); delete e; return nullptr; }
    for (int s = 0; s < 2; ++s) {
        relayout_bwt(e->idx.bwt[s], e->blocks[s]);
        DevBwt &d = e->bwt[s];
        d.blocks = reinterpret_cast<const uint4 *>(e->blocks[s].data());
        d.sa = e->idx.bwt[s].sa.data();
        d.primary = e->idx.bwt[s].primary; d.seq_len = e->idx.bwt[s].seq_len;
        for (int i = 0; i < 5; ++i) d.L2[i] = e->idx.bwt[s].L2[i];
        d.n_blocks = (uint32_t)e->blocks[s].size();
    }
    return e;
}
void emul_close(void *h) { delete (Emul *)h; }

void emul_occ4(void *h, int which, uint32_t k, uint32_t *out) { occ4(((Emul *)h)->bwt[which], k, out); }
uint32_t emul_sa(void *h, int which, uint32_t k) { return sa_lookup(((Emul *)h)->bwt[which], k); }

// widths + search for n reads; fwd = nt4 codes, read orientation, `stride` bytes per read
int emul_align(void *h, const fqb_gap_opt_t *gopt, int n, int stride, const uint8_t *fwd, const int32_t *lens,
               int arena_cap, int out_cap, fqb_aln_t *out, int32_t *n_aln, int32_t *status, uint32_t *pops_occ) {
    Emul *e = (Emul *)h;
    int max_len = 0;
    for (int r = 0; r < n; ++r) if (lens[r] > max_len) max_len = lens[r];
    SearchOpt so = make_search_opt(*gopt, max_len);
    if (so.n_buckets > 128) return -1;
    int32_t maxdiff[FQB_MAX_READ_LEN + 1];
    fill_maxdiff_table(*gopt, maxdiff);
    std::vector<uint4> arena((size_t)arena_cap);
    std::vector<uint32_t> heads((size_t)so.n_buckets);
    std::vector<uint16_t> heads16((size_t)so.n_buckets);
    std::vector<uint32_t> w0(max_len + 1), w1(max_len + 1), s0(so.seed_len + 1), s1(so.seed_len + 1);
    for (int r = 0; r < n; ++r) {
        const uint8_t *f = fwd + (size_t)r * stride;
        int len = lens[r];
        cal_width(e->bwt[0], f, len, 0, 0, len, w0.data());
        cal_width(e->bwt[1], f, len, 1, 0, len, w1.data());
        bool seeded = len > gopt->seed_len;
        if (seeded) {
            cal_width(e->bwt[0], f, len, 0, len - gopt->seed_len, gopt->seed_len, s0.data());
            cal_width(e->bwt[1], f, len, 1, len - gopt->seed_len, gopt->seed_len, s1.data());
        }
        SearchOpt lo = so;
        lo.seed_len = gopt->seed_len < len ? gopt->seed_len : 0x7fffffff;
        auto run = [&](auto &lane) {
            lane.bwt = e->bwt; lane.opt = &lo; lane.fwd = f;
            lane.w[0] = w0.data(); lane.w[1] = w1.data();
            lane.sw[0] = seeded ? s0.data() : nullptr; lane.sw[1] = seeded ? s1.data() : nullptr;
            lane.arena = arena.data(); lane.arena_cap = (uint32_t)arena_cap;
            lane.head_stride = 1;
            lane.out = reinterpret_cast<Hit *>(out + (size_t)r * out_cap); lane.out_cap = out_cap;
            int n_ambig = 0;
            for (int j = 0; j < len; ++j) n_ambig += f[j] > 3;
            lane.st_iter = lane.st_mempop = lane.st_skip = lane.st_exact = lane.st_expand = lane.st_push = lane.st_hit = lane.st_adiff = lane.st_gapok = lane.st_am = 0; for (auto &x : lane.st_x) x = 0; lane.st_run = 0;
            LaneStatus st = lane.begin(len, maxdiff[len], n_ambig);
            while (st == kLaneRunning || st == kLaneHit) {
                if (st == kLaneHit) lane.shadow_serial();
                st = lane.step();
            }
            n_aln[r] = lane.n_aln; status[r] = (int32_t)st;
            g_stats[0] += lane.st_iter; g_stats[1] += lane.st_mempop; g_stats[2] += lane.st_skip; g_stats[3] += lane.st_exact;
            g_stats[4] += lane.st_expand; g_stats[5] += lane.st_push; g_stats[6] += lane.st_hit; g_stats[7] += lane.top; g_stats[8] += 1; g_stats[9] += lane.st_adiff; g_stats[10] += lane.st_gapok; g_stats[11] += lane.st_am; for (int q = 0; q < 20; ++q) g_stats[12 + q] += lane.st_x[q];
            if (pops_occ) { pops_occ[2 * r] = lane.n_pops; pops_occ[2 * r + 1] = lane.n_occ; }
            g_stage[0] += lane.st_stage_hit; g_stage[1] += lane.st_stage_miss; g_stage[2] += lane.st_stage_stale;
        };
        if (arena_cap < 65535 && (g_search_var & 1)) { SearchLane<uint16_t, false, 5> lane; lane.heads = heads16.data(); run(lane); }
        else if (arena_cap < 65535) { SearchLane<uint16_t, false> lane; lane.heads = heads16.data(); run(lane); }
        else { SearchLane<uint32_t, true> lane; lane.heads = heads.data(); run(lane); }
    }
    return 0;
}


// bwt_cal_width's total lower bound per read and strand (the search queue is ordered by the smaller one)
extern "C" int emul_width_bid(void *h, int n, int stride, const uint8_t *fwd, const int32_t *lens, int32_t *bid2) {
    Emul *e = (Emul *)h;
    std::vector<uint32_t> w(FQB_MAX_READ_LEN + 2);
    for (int r = 0; r < n; ++r)
        for (int a = 0; a < 2; ++a) {
            cal_width(e->bwt[a], fwd + (size_t)r * stride, lens[r], a, 0, lens[r], w.data());
            bid2[2 * r + a] = width_bid(w[lens[r] - 1]);
        }
    return 0;
}

// Paired-end resolution with the SAME decomposition the kernels use (provisional draw counts -> prefix sum ->
// sequential pass over multi-interval reads -> per-read jump-ahead; histogram -> host infer_isize -> pair_one).
int emul_pe_batch(void *h, const fqb_gap_opt_t *gopt, const fqb_pe_opt_t *popt, int n_pairs, const int32_t *len,
                  const int32_t *full_len, const uint8_t *filtered, const int32_t *n_aln, const fqb_aln_t *aln, int aln_cap,
                  uint64_t *rng_calls, fqb_isize_t *last_ii, fqb_read_t *rows, fqb_isize_t *ii_out) {
    Emul *e = (Emul *)h;
    const int n = 2 * n_pairs;
    int32_t maxdiff[FQB_MAX_READ_LEN + 1], g_log_n[256];
    fill_maxdiff_table(*gopt, maxdiff);
    fill_log_n(g_log_n);
    const uint64_t x0 = lcg_seed(e->idx.seed);
    auto hits = [&](int r) { return reinterpret_cast<const Hit *>(aln + (size_t)r * aln_cap); };
    auto nh = [&](int r) { return filtered[r] ? 0 : n_aln[r]; };
    std::vector<uint64_t> guess(n), scanned(n), cum_extra;
    std::vector<uint32_t> multi;
    uint64_t run = 0;
    for (int r = 0; r < n; ++r) {
        int na = nh(r), nb = na ? count_best(hits(r), na) : 0;
        guess[r] = nb == 1 ? 2 : 0;
        if (nb > 1) multi.push_back((uint32_t)r);
        scanned[r] = run; run += guess[r];
    }
    uint64_t extra = 0;
    for (uint32_t r : multi) {
        fqb_read_t tmp; memset(&tmp, 0, sizeof tmp);
        extra += se_choose(hits(r), nh(r), lcg_advance(x0, *rng_calls + scanned[r] + extra), tmp);
        cum_extra.push_back(extra);
    }
    for (int r = 0; r < n; ++r) {
        fqb_read_t &row = rows[r];
        memset(&row, 0, sizeof row);
        row.len = len[r]; row.full_len = full_len[r]; row.clip_len = len[r]; row.filtered = filtered[r];
        row.extra_flag = kSamPaired | ((r & 1) ? kSamRead2 : kSamRead1);
        int na = nh(r);
        row.n_aln = (uint16_t)na;
        if (!na) continue;
        size_t lo = std::lower_bound(multi.begin(), multi.end(), (uint32_t)r) - multi.begin();
        uint64_t ex = lo ? cum_extra[lo - 1] : 0;
        se_choose(hits(r), na, lcg_advance(x0, *rng_calls + scanned[r] + ex), row);
        row.pos = hit_position(e->bwt, row.strand, row.sa, row.len);
        row.seQ = row.mapQ = (uint8_t)approx_mapq(row.c1, row.c2, row.n_mm, maxdiff[row.len], g_log_n);
    }
    *rng_calls += run + extra;
    std::vector<uint32_t> hist(kIsizeBins, 0);
    int max_len = 1;
    for (int p = 0; p < n_pairs; ++p) {
        const fqb_read_t &a = rows[2 * p], &b = rows[2 * p + 1];
        if (a.mapQ >= 20 && b.mapQ >= 20) {
            uint64_t x = a.pos < b.pos ? (uint64_t)b.pos + b.len - a.pos : (uint64_t)a.pos + a.len - b.pos;
            if (x < 100000) ++hist[x];
        }
        if (a.len > max_len) max_len = a.len;
        if (b.len > max_len) max_len = b.len;
    }
    fqb_isize_t ii;
    infer_isize_hist(hist.data(), max_len, popt->ap_prior, (int64_t)e->idx.bwt[0].seq_len, ii);
    if (ii.avg < 0.0 && last_ii->avg > 0.0) ii = *last_ii;
    if (popt->force_isize) { ii.low = ii.high = 0; ii.avg = ii.std = -1.0; }
    std::vector<int32_t> pen;
    fill_isize_penalty(ii, pen);
    PairParams pp;
    pp.high = ii.high; pp.high_bayesian = ii.high_bayesian; pp.max_isize = popt->max_isize; pp.s_mm = gopt->s_mm;
    pp.max_occ = popt->max_occ; pp.n_multi = popt->n_multi; pp.N_multi = popt->N_multi;
    pp.penalty = pen.data(); pp.g_log_n = g_log_n; pp.sw_on = 0;
    std::vector<uint64_t> arr(8192);
    for (int p = 0; p < n_pairs; ++p)
        if (!pair_one(e->bwt, &rows[2 * p], &rows[2 * p + 1], hits(2 * p), nh(2 * p), hits(2 * p + 1), nh(2 * p + 1), pp, arr.data(), kPairArrCap))
            if (!pair_one(e->bwt, &rows[2 * p], &rows[2 * p + 1], hits(2 * p), nh(2 * p), hits(2 * p + 1), nh(2 * p + 1), pp, arr.data(), 8192)) return -2;
    *ii_out = ii; *last_ii = ii;
    return 0;
}


// bwa_paired_sw + bwa_refine_gapped with the device functions (stride-1 scratch)
int emul_sw_refine(void *h, const fqb_pe_opt_t *popt, int n_pairs, int stride, const uint8_t *codes, const fqb_isize_t *ii,
                   fqb_read_t *rows, fqb_read_t *rows_after_sw) {
    Emul *e = (Emul *)h;
    const int64_t l_pac = e->idx.l_pac;
    const uint8_t *pac = e->idx.pac.data();
    std::vector<int32_t> ints(6 * 1100);
    std::vector<uint8_t> bytes(400000);
    DpScratch sc; sc.ints = ints.data(); sc.n_ints = (int)ints.size(); sc.bytes = bytes.data(); sc.n_bytes = (int)bytes.size(); sc.istride = sc.bstride = 1;
    if (popt->is_sw && ii->avg >= 0.0) {
        SwParams sp;
        sp.avg = ii->avg; sp.std = ii->std; sp.l_pac = l_pac;
        sp.s_old_add = -4.343 * std::log(ii->ap_prior / l_pac);
        sp.s_new_add = (int)(-4.343 * std::log(.5 * std::erfc(M_SQRT1_2 * 1.5) + .499));
        for (int p = 0; p < n_pairs; ++p) {
            fqb_read_t *p0 = rows + 2 * p, *p1 = p0 + 1;
            if (p0->filtered) { if (p1->filtered) continue; p0->filtered = 0; }
            else if (p1->filtered) p1->filtered = 0;
            if ((p0->mapQ >= 17 || p1->mapQ >= 17) && (p0->extra_flag & kSamProper) == 0)
                if (!paired_sw_one(pac, p0, p1, codes + (size_t)(2 * p) * stride, codes + (size_t)(2 * p + 1) * stride, sp, sc)) return -1;
        }
    }
    if (rows_after_sw) memcpy(rows_after_sw, rows, sizeof(fqb_read_t) * 2 * (size_t)n_pairs);
    for (int r = 0; r < 2 * n_pairs; ++r) {
        fqb_read_t &s = rows[r];
        ReadSeq Q; Q.fwd = codes + (size_t)r * stride; Q.len = s.len; Q.strand = s.strand;
        if (!s.filtered && !(s.type == kTypeNoMatch || s.type == kTypeMateSW || s.n_gapo == 0)) {
            int nc = refine_gapped(l_pac, pac, Q, &s.pos, (s.strand ? 1 : -1) * (s.n_gapo + s.n_gape), s.cigar, FQB_MAX_CIGAR, sc);
            if (nc < 0) return -2;
            s.n_cigar = (uint8_t)nc; s.has_cigar = 1;
        }
        if (s.type != kTypeNoMatch) s.nm = (uint16_t)cal_nm(s, Q, l_pac, pac);
        correct_trimmed(s);
    }
    return 0;
}

}  // extern "C"

// Row a12 without a GPU: the device function classify_pair (StatCollector::AddAlignment / ProcessPairStatus) over final
// rows, pair by pair in file order with host-side accumulators, and the product's own InsertSizeTable formatter on the
// PairStat it leaves.  State (insert-size histogram, duplicate keys, contig counters) persists across batches.
struct EmulStats {
    StatsTables T;
    std::vector<uint32_t> contig_ctr, contig_first;
    std::vector<unsigned long long> isize_dist, scalars, dup_keys;
    unsigned long long dup_count = 0;
    StatAccum A;
    unsigned long long fsc[5] = {0, 0, 0, 0, 0};     // both_filtered, both_unmapped, low_mapq, retained, bases
    std::vector<uint32_t> depth, q20, q30;            // [n_sites]
    std::vector<unsigned long long> emp;              // [4][256]
    std::vector<PileupColumn> pileup;                 // [n_markers]
    unsigned long long n_reads = 0;
};

extern "C" {

void *emul_stats_open(void *h, const char *index_prefix, const fqb_gap_opt_t *gopt, char *err, int errlen) {
    Emul *e = (Emul *)h;
    EmulStats *s = new EmulStats();
    std::string msg;
    if (!build_stats_tables(e->idx, index_prefix, *gopt, "", s->T, msg)) { strncpy(err, msg.c_str(), errlen - 1); delete s; return nullptr; }
    const size_t nc = s->T.contigs.size();
    s->contig_ctr.assign(nc * 4, 0); s->contig_first.assign(nc, 0xffffffffu);
    s->isize_dist.assign(4096, 0); s->scalars.assign(16, 0); s->dup_keys.assign(1u << 16, 0);
    s->A.contig_ctr = s->contig_ctr.data(); s->A.contig_first = s->contig_first.data();
    s->A.isize_dist = s->isize_dist.data(); s->A.scalars = s->scalars.data();
    s->A.dup_keys = s->dup_keys.data(); s->A.dup_cap = (uint32_t)s->dup_keys.size(); s->A.dup_count = &s->dup_count;
    const size_t ns = s->T.n_sites ? s->T.n_sites : 1;
    s->depth.assign(ns, 0); s->q20.assign(ns, 0); s->q30.assign(ns, 0); s->emp.assign(4 * 256, 0);
    s->pileup.assign(s->T.markers.size(), PileupColumn());
    return s;
}
void emul_stats_close(void *st) { delete (EmulStats *)st; }

// rows: final rows of one batch (r = 2 * pair + end), modified as AddAlignment modifies them (bridge-check demotions).
// Appends the batch's InsertSizeTable lines to out; returns their length, or -1 if cap is too small.
long long emul_stats_batch(void *st, int n_pairs, unsigned long long pair_base, int cal_dup, fqb_read_t *rows, uint8_t *add_out, char *out, long long cap) {
    EmulStats *s = (EmulStats *)st;
    std::string text;
    char name[32];
    for (int p = 0; p < n_pairs; ++p) {
        PairStat o;
        classify_pair(s->T.contigs.data(), (int)s->T.contigs.size(), rows[2 * p], rows[2 * p + 1], (uint32_t)(pair_base + p), cal_dup, s->A, o);
        s->fsc[0] += o.both_filtered; s->fsc[1] += o.both_unmapped; s->fsc[2] += o.low_mapq; s->fsc[3] += o.retained;
        s->fsc[4] += (unsigned long long)(rows[2 * p].full_len + rows[2 * p + 1].full_len);
        if (add_out) { add_out[2 * p] = o.add[0]; add_out[2 * p + 1] = o.add[1]; }
        if (o.line_kind == 0) continue;
        snprintf(name, sizeof name, "r%011llu", pair_base + p);
        append_isize_line(s->T, o, rows[2 * p], rows[2 * p + 1], name, text);
    }
    if ((long long)text.size() > cap) return -1;
    memcpy(out, text.data(), text.size());
    return (long long)text.size();
}
void emul_stats_totals(void *st, unsigned long long *isize_dist /*4096*/, unsigned long long *fsc /*5*/, unsigned long long *scalars /*16*/, unsigned long long *dup_count) {
    EmulStats *s = (EmulStats *)st;
    memcpy(isize_dist, s->isize_dist.data(), 4096 * 8);
    memcpy(fsc, s->fsc, sizeof s->fsc);
    memcpy(scalars, s->scalars.data(), 16 * 8);
    *dup_count = s->dup_count;
}

}  // extern "C"

// Rows a13 / a14 without a GPU.  emul_stats_bases is a SERIAL RESTATEMENT of bases_kernel's per-base walk
// (fq_stats_kernels.cu; AddSingleAlignment, src/StatCollector.cpp:424-621) - reads in file order, offsets in order, which
// is the arrival order the kernel's tuple keys sort back into - so what it pins is everything around that loop: the
// add[] decisions of classify_pair, the side tables of build_stats_tables and the writers of write_summary_files.
extern "C" {

int emul_stats_bases(void *st, void *h, int n_reads, int stride, const uint8_t *codes /*nt4, as sequenced*/, const uint8_t *quals /*ASCII*/,
                     const fqb_read_t *rows, const uint8_t *add) {
    EmulStats *s = (EmulStats *)st;
    Emul *e = (Emul *)h;
    const uint8_t *pac = e->idx.pac.data();
    for (int r = 0; r < n_reads; ++r) {
        if (!add[r]) continue;
        const fqb_read_t &row = rows[r];
        const uint8_t *fwd = codes + (size_t)r * stride, *ql = quals + (size_t)r * stride;
        const int full = row.full_len;
        for (int off = 0; off < full; ++off) {
            bool is_m = false;
            uint32_t x = 0;
            if (row.has_cigar) {
                int y = 0; uint32_t xx = row.pos;
                for (int k = 0; k < row.n_cigar; ++k) {
                    const int op = row.cigar[k] >> 14, cl = row.cigar[k] & 0x3fff;
                    if (op == kOpM) { if (off < y + cl) { is_m = true; x = xx + (uint32_t)(off - y); break; } y += cl; xx += cl; }
                    else if (op == kOpD) xx += cl;
                    else { if (off < y + cl) break; y += cl; }
                }
            } else { is_m = off < row.len; x = row.pos + (uint32_t)off; }
            if (!is_m) continue;
            const int fo = row.strand ? full - 1 - off : off;
            uint32_t rb = fwd[fo]; if (row.strand && rb < 4) rb = 3 - rb;
            const uint32_t q = (uint32_t)ql[fo] - 33u;
            const int cycle = fo;
            const uint32_t site = s->T.site[x];
            if (site & kSiteMarker) {
                PileupColumn &c = s->pileup[(size_t)s->T.marker_at[x]];
                c.seq.push_back("ACGTN"[rb > 4 ? 4 : rb]); c.qual.push_back((char)q); c.cycle.push_back(cycle);
                c.maq.push_back((unsigned char)(row.mapQ + 33)); c.strand.push_back(row.strand != 0);
            }
            if ((site & kSiteMask) == kSiteNone) continue;
            const uint32_t sid = site & kSiteMask;
            ++s->depth[sid];
            if ((int8_t)q >= 20) { ++s->q20[sid]; if ((int8_t)q >= 30) ++s->q30[sid]; }
            const uint32_t refb = pac[x >> 2] >> ((~x & 3) << 1) & 3;
            const bool mis = rb < 4 && refb != rb && !(site & kSiteDbsnp);
            ++s->emp[0 * 256 + (q & 255)]; ++s->emp[2 * 256 + ((uint32_t)cycle & 255)];
            if (mis) { ++s->emp[1 * 256 + (q & 255)]; ++s->emp[3 * 256 + ((uint32_t)cycle & 255)]; }
        }
    }
    s->n_reads += (unsigned long long)n_reads;
    return 0;
}

// StatsTotals from the accumulators, then the product's writers (the InsertSizeTable must already be at <prefix>.InsertSizeTable)
int emul_stats_finish(void *st, const fqb_gap_opt_t *gopt, const char *out_prefix, const char *fq1, const char *fq2, char *err, int errlen) {
    EmulStats *s = (EmulStats *)st;
    StatsTotals S;
    S.depth = s->depth; S.q20 = s->q20; S.q30 = s->q30; S.emp = s->emp; S.isize_dist = s->isize_dist;
    S.num_pcr_dup = s->scalars[0]; S.num_pair_reads = s->scalars[1];
    S.contig_ctr = s->contig_ctr; S.contig_first = s->contig_first;
    S.pileup = s->pileup;
    FileCounters F;
    F.FileName1 = fq1; F.FileName2 = fq2;
    F.TotalFiltered = (long long)s->fsc[0]; F.BwaUnmapped = (long long)s->fsc[1]; F.TotalMAPQ = (long long)s->fsc[2];
    F.TotalRetained = (long long)s->fsc[3]; F.NumBase = (long long)s->fsc[4]; F.NumRead = (long long)s->n_reads;
    S.files.push_back(F);
    std::string msg;
    if (s->scalars[2]) { strncpy(err, "insert size out of range or duplicate table full", errlen - 1); return -1; }
    if (!write_summary_files(s->T, S, *gopt, out_prefix, msg)) { strncpy(err, msg.c_str(), errlen - 1); return -1; }
    return 0;
}

}  // extern "C"
