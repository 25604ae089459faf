"""Host halves of the hot path, no GPU needed (rows a8, a3, a6, a4's bucket sizing).

Row a8: the product's infer_isize over the device's insert-size histogram (fq_hostmath.cpp:
infer_isize_hist) against the C oracle's restatement of libbwa/bwape.c:49-117 over the same pairs as a sorted array
(oracle/fq_oracle_pe.c, itself pinned to the reference by the golden pair-stage rows).  Every field must be bit-equal:
avg / std / ap_prior feed integer decisions in pairing() and mate rescue."""
import ctypes as C
import math
import zlib

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

L_BWT = 13_217_394          # 2 x l_pac of the 10k-marker index; only enters through ap_prior / L


def _both(isizes, read_len=100, ap_prior=1e-5, mapq=None):
    isizes = np.asarray(isizes, np.int64)
    n = len(isizes)
    rows = np.zeros(2 * n, _abi.READ_DTYPE)
    rows["len"] = read_len
    rows["pos"][0::2] = 5000
    rows["pos"][1::2] = 5000 + isizes - read_len          # x = p1.pos + p1.len - p0.pos = isize (isize > read_len)
    rows["mapQ"] = 37 if mapq is None else np.repeat(mapq, 2)
    orc = fx.build_oracle()
    orc.orc_infer_isize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int64]
    want = _abi.ISize()
    rc_o = orc.orc_infer_isize(n, rows.ctypes.data, C.byref(want), ap_prior, L_BWT)
    keep = isizes if mapq is None else isizes[np.asarray(mapq) >= 20]
    hist = np.bincount(keep[keep < 100000], minlength=100000).astype(np.uint32)
    lib = fx.host_lib()
    lib.fqb_infer_isize_hist.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int64, C.c_void_p]
    got = _abi.ISize()
    rc_p = lib.fqb_infer_isize_hist(hist.ctypes.data, read_len, ap_prior, L_BWT, C.byref(got))
    return rc_o, want, rc_p, got


def _same(a, b):
    for f in ("low", "high", "high_bayesian"):
        assert getattr(a, f) == getattr(b, f), f
    for f in ("avg", "std", "ap_prior"):
        x, y = getattr(a, f), getattr(b, f)
        assert x == y or (math.isnan(x) and math.isnan(y)), (f, x, y)


CASES = {
    "normal_300_30": lambda r: r.normal(300, 30, 5000),
    "normal_500_80": lambda r: r.normal(500, 80, 262144),
    "tight": lambda r: r.normal(250, 2, 400),
    "twenty": lambda r: r.normal(300, 30, 20),
    "outliers": lambda r: np.concatenate([r.normal(350, 40, 3000), r.uniform(2000, 99999, 200)]),
    "bimodal": lambda r: np.concatenate([r.normal(200, 15, 2000), r.normal(900, 50, 2000)]),
    "beyond_100k": lambda r: np.concatenate([r.normal(300, 30, 500), r.uniform(100000, 400000, 300)]),
    "constant": lambda r: np.full(100, 321.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_infer_isize_hist_matches_oracle(name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    isizes = np.maximum(CASES[name](rng).astype(np.int64), 101)
    rc_o, want, rc_p, got = _both(isizes)
    assert (rc_o == 0) == (rc_p == 1)
    _same(want, got)
    if name != "constant": assert rc_p == 1 and got.low <= got.avg <= got.high <= got.high_bayesian


def test_infer_isize_hist_failure_and_mapq_filter():
    rng = np.random.default_rng(4)
    rc_o, want, rc_p, got = _both(np.maximum(rng.normal(300, 30, 19).astype(np.int64), 101))      # fewer than 20 pairs
    assert rc_o == -1 and rc_p == 0 and got.avg == -1.0 and got.std == -1.0 and got.high == 0
    isizes = np.maximum(rng.normal(300, 30, 4000).astype(np.int64), 101)
    mapq = rng.choice([0, 10, 19, 20, 25, 37], 4000)                                               # only pairs with mapQ >= 20 count
    rc_o, want, rc_p, got = _both(isizes, mapq=mapq)
    assert rc_o == 0 and rc_p == 1
    _same(want, got)


def test_isize_penalty_table():
    """fill_isize_penalty = the expression in __pairing_aux's macro (libbwa/bwape.h:62) for every insert size up to
    high_bayesian, evaluated with the same libm."""
    rng = np.random.default_rng(6)
    _, _, rc, ii = _both(np.maximum(rng.normal(300, 30, 5000).astype(np.int64), 101))
    assert rc == 1
    lib = fx.host_lib()
    lib.fqb_isize_penalty.restype = C.c_int64
    lib.fqb_isize_penalty.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    n = lib.fqb_isize_penalty(C.byref(ii), None, 0)
    assert n == ii.high_bayesian + 1
    t = np.zeros(n, np.int32)
    assert lib.fqb_isize_penalty(C.byref(ii), t.ctypes.data, n) == n
    libm = C.CDLL("libm.so.6")
    libm.erfc.restype = C.c_double; libm.erfc.argtypes = [C.c_double]
    libm.log.restype = C.c_double; libm.log.argtypes = [C.c_double]
    for l in list(range(0, n, 7)) + [n - 1, int(ii.avg)]:
        v = -4.343 * libm.log(.5 * libm.erfc(math.sqrt(0.5) * abs(l - ii.avg) / ii.std)) + .499
        assert t[l] == int(v), l                       # C's (int) truncates toward zero, as Python's int() does
    assert t[int(ii.avg)] == 3 and t[0] > t[int(ii.avg) - 60] > t[int(ii.avg)]


def test_maxdiff_and_log_n_tables_match_oracle():
    """Rows a3 / a6: the tables fqb_create ships to the device against the oracle's bwa_cal_maxdiff (pinned to the
    reference's in test_oracle_vs_ref.py) and g_log_n."""
    lib, orc = fx.host_lib(), fx.build_oracle()
    orc.orc_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    lib.fqb_host_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    md = np.zeros(257, np.int32); ln = np.zeros(256, np.int32)
    for fnr in (0.04, 0.02, 0.001):                              # 0.02 is FASTQuick's (max_diff 5 / 7 at 100 / 150 bp)
        g.fnr = fnr
        assert lib.fqb_host_tables(C.byref(g), md.ctypes.data, ln.ctypes.data) == 0
        assert [int(v) for v in md] == [orc.orc_cal_maxdiff(l, 0.02, fnr) for l in range(257)]
    assert md[100] == 8 and md[150] == 10                        # fnr 0.001
    g.fnr = 0.04
    lib.fqb_host_tables(C.byref(g), md.ctypes.data, ln.ctypes.data)
    assert md[100] == 5 and md[150] == 6                         # bwa aln's documented values at its default fnr
    g.fnr = -1.0; g.max_diff = 3
    lib.fqb_host_tables(C.byref(g), md.ctypes.data, ln.ctypes.data)
    assert (md == 3).all()
    want = (C.c_int * 256)()
    orc.orc_fill_log_n(want)
    assert list(ln) == list(want) and ln[1] == 0 and ln[2] == 3 and ln[255] == 24


def test_search_bucket_count():
    """gap_init_stack: 3 (max_diff + 1) + 11 * 2 + 4 * 7 = 68 / 74 score buckets for 100 / 150 bp at FASTQuick's fnr 0.02
    (SURVEY 8); max_gapo is clamped to max_diff when -o exceeds it."""
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.fnr = 0.02
    assert lib.fqb_search_buckets(C.byref(g), 100) == 68 and lib.fqb_search_buckets(C.byref(g), 150) == 74
    assert lib.fqb_search_buckets(C.byref(g), 10) == (1 + 1) * 3 + (1 + 1) * 11 + 7 * 4      # max_diff 1
    g.max_gapo = 3
    assert lib.fqb_search_buckets(C.byref(g), 100) == 6 * 3 + 4 * 11 + 7 * 4
    assert lib.fqb_search_buckets(C.byref(g), 10) == 2 * 3 + 2 * 11 + 7 * 4                  # -o 3 clamped to max_diff 1
    assert lib.fqb_search_buckets(C.byref(g), 1000) < 0


def test_drand48_zero_draw_index():
    """The draw at which drand48() returns exactly 0.0 (the one case the device's draw count does not follow,
    libbwa/bwase.c:33-36): found by fqb_drand48_zero_index, checked by stepping the generator around it."""
    import ctypes as C
    lib = fx.host_lib()
    lib.fqb_drand48_zero_index.restype = C.c_uint64
    A, Cc, M = 0x5DEECE66D, 0xB, (1 << 48) - 1

    def advance(x, n):
        a, c, aa, cc = A, Cc, 1, 0
        while n:
            if n & 1:
                aa, cc = (aa * a) & M, (cc * a + c) & M
            c, a, n = ((a + 1) * c) & M, (a * a) & M, n >> 1
        return (aa * x + cc) & M

    for seed in (11, 0, 1, 12345, 0xffffffff):
        n = int(lib.fqb_drand48_zero_index(C.c_uint32(seed)))
        x0 = (seed << 16) | 0x330E
        assert 1 <= n <= 1 << 48
        assert advance(x0, n) == 0                      # drand48() == 0.0 at call n ...
        assert advance(x0, n - 1) != 0 or n == 1 << 48  # ... and the closed form is not off by one
    assert int(lib.fqb_drand48_zero_index(C.c_uint32(11))) == 79023531276618      # bns->seed of every BWA index
    # brute force for a state that reaches zero soon: seed state X with a X + c = 0 (mod 2^48) is one step away
    x = (-(Cc) * pow(A, -1, 1 << 48)) & M
    if x & 0xffff == 0x330E and x >> 16 < 1 << 32:
        assert int(lib.fqb_drand48_zero_index(C.c_uint32(x >> 16))) == 1
