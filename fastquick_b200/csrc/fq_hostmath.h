// Host-side libm-dependent tables.  Integer decisions on the path that go through
// glibc exp/log/erfc in the reference are evaluated on the host with the same
// libm and shipped to the device as tables/scalars.
#pragma once
#include <cstdint>
#include <vector>
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
// infer_isize (libbwa/bwape.c:49-117) from a histogram of the insert sizes < 100000 of the pairs whose
// ends both have mapQ >= 20 (the device collects it); same arithmetic, same summation order as the
// reference's pass over the sorted array.  Returns false when inference fails (ii = "unset").
constexpr int kIsizeBins = 100000;
bool infer_isize_hist(const uint32_t *hist, int max_len, double ap_prior, int64_t L, fqb_isize_t &ii);
// penalty[l] = (int)(-4.343*log(.5*erfc(M_SQRT1_2*fabs(l-avg)/std))+.499) for l = 0..high_bayesian (libbwa/bwape.h:62)
void fill_isize_penalty(const fqb_isize_t &ii, std::vector<int32_t> &table);
}  // namespace fqb
