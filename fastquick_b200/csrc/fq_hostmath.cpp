#include "fq_hostmath.h"
#include "fq_common.h"
#include <cmath>

namespace fqb {

int cal_maxdiff(int l, double err, double thres) {
    double elambda = std::exp(-l * err), sum = elambda, y = 1.0;
    unsigned x = 1;
    for (int k = 1; k < 1000; ++k) {
        y *= l * err;
        x *= (unsigned)k;                     // the reference multiplies an int that wraps the same way
        sum += elambda * y / (int)x;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

void fill_maxdiff_table(const fqb_gap_opt_t &o, int32_t *t) {
    for (int l = 0; l <= FQB_MAX_READ_LEN; ++l) t[l] = o.fnr > 0.0 ? cal_maxdiff(l, 0.02 /*BWA_AVG_ERR*/, o.fnr) : o.max_diff;
}

void fill_log_n(int32_t *g) {
    g[0] = 0;
    for (int i = 1; i < 256; ++i) g[i] = (int)(4.343 * std::log(i) + 0.5);
}

SearchOpt make_search_opt(const fqb_gap_opt_t &o, int max_len) {
    SearchOpt s;
    s.s_mm = o.s_mm; s.s_gapo = o.s_gapo; s.s_gape = o.s_gape; s.mode = o.mode;
    s.indel_end_skip = o.indel_end_skip; s.max_del_occ = o.max_del_occ; s.max_entries = o.max_entries;
    int md = o.fnr > 0.0 ? cal_maxdiff(max_len, 0.02, o.fnr) : o.max_diff;
    s.max_gapo = md < o.max_gapo ? md : o.max_gapo;
    s.max_gape = o.max_gape;
    s.max_seed_diff = o.max_seed_diff; s.seed_len = o.seed_len; s.max_top2 = o.max_top2;
    s.n_buckets = (md + 1) * o.s_mm + (s.max_gapo + 1) * o.s_gapo + (o.max_gape + 1) * o.s_gape;   // gap_init_stack
    return s;
}

bool infer_isize_hist(const uint32_t *hist, int max_len_in, double ap_prior, int64_t L, fqb_isize_t &ii) {
    ii.avg = ii.std = -1.0; ii.ap_prior = 0.0;
    ii.low = ii.high = ii.high_bayesian = 0; ii.pad_ = 0;
    uint64_t tot = 0;
    for (int v = 0; v < kIsizeBins; ++v) tot += hist[v];
    if (tot < 20) return false;
    auto at = [&](uint64_t idx) {           // value at position idx of the sorted array
        uint64_t c = 0;
        for (int v = 0; v < kIsizeBins; ++v) { c += hist[v]; if (c > idx) return v; }
        return kIsizeBins - 1;
    };
    const int itot = (int)tot;
    int p25 = at((uint64_t)(int)(itot * 0.25 + 0.5));
    int p75 = at((uint64_t)(int)(itot * 0.75 + 0.5));
    int max_len = max_len_in < 1 ? 1 : max_len_in;
    int tmp = (int)(p25 - 2.0 * (p75 - p25) + .499);
    ii.low = tmp > max_len ? (uint32_t)tmp : (uint32_t)max_len;
    ii.high = (uint32_t)(int)(p75 + 2.0 * (p75 - p25) + .499);
    uint64_t x = 0;
    int n = 0;
    for (uint32_t v = ii.low; v <= ii.high && v < (uint32_t)kIsizeBins; ++v) { n += (int)hist[v]; x += (uint64_t)v * hist[v]; }
    ii.avg = (double)x / n;
    for (uint32_t v = ii.low; v <= ii.high && v < (uint32_t)kIsizeBins; ++v) {
        const double t = ((uint64_t)v - ii.avg) * ((uint64_t)v - ii.avg);
        for (uint32_t c = 0; c < hist[v]; ++c) ii.std += t;       // one addition per pair, in sorted order, as the reference does
    }
    ii.std = std::sqrt(ii.std / n);
    double y;
    for (y = 1.0; y < 10.0; y += 0.01)
        if (.5 * std::erfc(y / M_SQRT2) < ap_prior / L * (y * ii.std + ii.avg)) break;
    ii.high_bayesian = (uint32_t)(y * ii.std + ii.avg + .499);
    uint64_t n_ap = 0;
    for (int v = 0; v < kIsizeBins; ++v) if ((uint32_t)v > ii.high_bayesian) n_ap += hist[v];
    ii.ap_prior = .01 * (n_ap + .01) / itot;
    if (ii.ap_prior < ap_prior) ii.ap_prior = ap_prior;
    if (std::isnan(ii.std) || p75 > 100000) {
        ii.low = ii.high = ii.high_bayesian = 0; ii.avg = ii.std = -1.0;
        return false;
    }
    for (y = 1.0; y < 10.0; y += 0.01)
        if (.5 * std::erfc(y / M_SQRT2) < ap_prior / L * (y * ii.std + ii.avg)) break;
    ii.high_bayesian = (uint32_t)(y * ii.std + ii.avg + .499);
    return true;
}

void fill_isize_penalty(const fqb_isize_t &ii, std::vector<int32_t> &t) {
    t.clear();
    if (ii.high == 0) return;
    t.resize((size_t)ii.high_bayesian + 1);
    for (uint32_t l = 0; l <= ii.high_bayesian; ++l)
        t[l] = (int32_t)(-4.343 * std::log(.5 * std::erfc(M_SQRT1_2 * std::fabs(l - ii.avg) / ii.std)) + .499);
}

}  // namespace fqb

// The host half of infer_isize and the pairing penalty table on their own (host only; the device supplies the histogram
// in the product path, fq_engine.cu: fqb_stage_pair).
extern "C" int fqb_infer_isize_hist(const uint32_t *hist, int32_t max_len, double ap_prior, int64_t L, fqb_isize_t *ii) {
    if (!hist || !ii || L <= 0) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    return fqb::infer_isize_hist(hist, max_len, ap_prior, L, *ii) ? 1 : 0;
}
extern "C" int64_t fqb_isize_penalty(const fqb_isize_t *ii, int32_t *out, int64_t cap) {
    if (!ii || (!out && cap > 0) || cap < 0) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    std::vector<int32_t> t;
    fqb::fill_isize_penalty(*ii, t);
    for (int64_t i = 0; i < (int64_t)t.size() && i < cap; ++i) out[i] = t[(size_t)i];
    return (int64_t)t.size();
}

// The libm-dependent tables the engine ships to the device (bwa_cal_maxdiff per read length, g_log_n), host only.
extern "C" int fqb_host_tables(const fqb_gap_opt_t *gopt, int32_t *maxdiff /*FQB_MAX_READ_LEN + 1*/, int32_t *log_n /*256*/) {
    if (!gopt || !maxdiff || !log_n) { fqb::set_error("null argument"); return FQB_ERR_ARG; }
    fqb::fill_maxdiff_table(*gopt, maxdiff);
    fqb::fill_log_n(log_n);
    return FQB_OK;
}

// gap_init_stack's bucket count (libbwa/bwtgap.c:18) for a batch whose longest read has max_len bases, with the max_gapo
// clamp of src/BwtMapper.cpp:73-81 applied: what sizes the per-read score-bucket heads in search_kernel.
extern "C" int fqb_search_buckets(const fqb_gap_opt_t *gopt, int32_t max_len) {
    if (!gopt || max_len < 0 || max_len > FQB_MAX_READ_LEN) { fqb::set_error("bad argument"); return FQB_ERR_ARG; }
    return fqb::make_search_opt(*gopt, max_len).n_buckets;
}
