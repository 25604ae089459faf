#include "fq_hostmath.h"
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

}  // namespace fqb
