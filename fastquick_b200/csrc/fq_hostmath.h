// Host-side libm-dependent tables.  Integer decisions on the path that go through
// glibc exp/log/erfc in the reference are evaluated on the host with the same
// libm and shipped to the device as tables/scalars.
#pragma once
#include <cstdint>
#include "../../include/fastquick_b200.h"
#include "fq_device_core.cuh"

namespace fqb {
// bwa_cal_maxdiff (libbwa/bwtaln.c:58-70)
int cal_maxdiff(int l, double err, double thres);
// max_diff for every read length 0..FQB_MAX_READ_LEN (fnr > 0 ? Poisson rule : opt->max_diff)
void fill_maxdiff_table(const fqb_gap_opt_t &o, int32_t *table /*FQB_MAX_READ_LEN+1*/);
// g_log_n (libbwa/bwase.c:602-606)
void fill_log_n(int32_t *g /*256*/);
// SearchOpt for a batch whose longest read is max_len (stack sizing + max_gapo clamp, src/BwtMapper.cpp:73-81)
SearchOpt make_search_opt(const fqb_gap_opt_t &o, int max_len);
}  // namespace fqb
