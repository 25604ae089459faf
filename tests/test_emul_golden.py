"""Kernel LOGIC of the pair / mate-rescue / refinement stages without a GPU (rows a6-a11): the device functions of
fq_device_pair.cuh and fq_device_dp.cuh, instantiated on the host by tests/emul with the decomposition the kernels use
(provisional drand48 draw counts -> prefix sum -> jump-ahead; insert-size histogram -> host infer_isize -> pair_one;
paired_sw_one; refine_gapped / cal_nm / correct_trimmed), fed with the REFERENCE's hit lists from the committed golden
vectors and compared with the reference's rows after every stage, over consecutive batches so that the RNG position and
last_ii carry over.  Needs neither a GPU nor oracle/_ref."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402
from test_golden import FIELDS, _case, _cigars_equal  # noqa: E402


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_emulated_pair_sw_refine_against_golden(name):
    small_index = make_golden.index_for(name)
    arrs, g, n, batch = _case(small_index, name)
    lib = fx.build_emul()
    lib.emul_open.restype = C.c_void_p
    err = C.create_string_buffer(256)
    h = C.c_void_p(lib.emul_open(small_index.prefix.encode(), err, 256))
    assert h, err.value
    host = fx.host_lib()
    gopt = _abi.GapOpt(); host.fqb_gap_opt_default(C.byref(gopt)); gopt.trim_qual = 15
    popt = _abi.PeOpt(); host.fqb_pe_opt_default(C.byref(popt))
    rng_calls = C.c_uint64(0)
    last_ii = _abi.ISize(); last_ii.avg = -1.0; last_ii.std = -1.0
    try:
        for b in range(n // batch):
            r0 = [g["b%d_e%d_rows0" % (b, e)] for e in (0, 1)]
            lens = np.zeros(2 * batch, np.int32); full = np.zeros(2 * batch, np.int32); filt = np.zeros(2 * batch, np.uint8)
            for e in (0, 1):
                lens[e::2] = r0[e]["len"]; full[e::2] = r0[e]["full_len"]; filt[e::2] = r0[e]["filtered"]
            cap = max(int(np.diff(g["b%d_e%d_aln_off" % (b, e)]).max()) for e in (0, 1))
            aln = np.zeros((2 * batch, cap), _abi.ALN_DTYPE); na = np.zeros(2 * batch, np.int32)
            for e in (0, 1):
                pad, cnt = fx.csr_to_padded(g["b%d_e%d_aln_off" % (b, e)], g["b%d_e%d_aln" % (b, e)], cap)
                aln[e::2] = pad; na[e::2] = cnt
            rows = np.zeros(2 * batch, _abi.READ_DTYPE)
            ii = _abi.ISize()
            rc = lib.emul_pe_batch(h, C.byref(gopt), C.byref(popt), batch, _abi.i32p(lens), _abi.i32p(full), _abi.u8p(filt), _abi.i32p(na),
                                   aln.ctypes.data_as(C.c_void_p), cap, C.byref(rng_calls), C.byref(last_ii),
                                   rows.ctypes.data_as(C.c_void_p), C.byref(ii))
            assert rc == 0
            gf, gi = g["b%d_isize_f" % b], g["b%d_isize_i" % b]
            assert (ii.avg, ii.std, ii.ap_prior) == tuple(gf) and (ii.low, ii.high, ii.high_bayesian) == tuple(int(x) for x in gi)
            for e in (0, 1):
                for f in FIELDS:
                    np.testing.assert_array_equal(rows[e::2][f], g["b%d_e%d_rows1" % (b, e)][f], err_msg="pe b%d e%d %s" % (b, e, f))
            sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
            rl = sub[0].shape[1]
            codes = np.zeros((2 * batch, rl), np.uint8)
            codes[0::2] = fx.NT4[sub[0]]; codes[1::2] = fx.NT4[sub[2]]
            codes = np.ascontiguousarray(codes)
            after_sw = np.zeros_like(rows)
            rc = lib.emul_sw_refine(h, C.byref(popt), batch, rl, _abi.u8p(codes), C.byref(ii), rows.ctypes.data_as(C.c_void_p),
                                    after_sw.ctypes.data_as(C.c_void_p))
            assert rc == 0
            for e in (0, 1):
                for f in FIELDS:
                    np.testing.assert_array_equal(after_sw[e::2][f], g["b%d_e%d_rows2" % (b, e)][f], err_msg="sw b%d e%d %s" % (b, e, f))
                _cigars_equal(rows[e::2], g["b%d_e%d_rows3" % (b, e)], "refine b%d e%d" % (b, e))
        assert rng_calls.value > 0
    finally:
        lib.emul_close.argtypes = [C.c_void_p]
        lib.emul_close(h)


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_emulated_pair_classification_against_golden(name):
    """Row a12: the device function classify_pair (AddAlignment / ProcessPairStatus) over the reference's final rows, then
    the product's InsertSizeTable formatter: the lines must be the reference's InsertSizeTable, the insert-size histogram
    its RawInsertSizeDist."""
    small_index = make_golden.index_for(name)
    arrs, g, n, batch = _case(small_index, name)
    lib = fx.build_emul()
    lib.emul_open.restype = C.c_void_p
    lib.emul_stats_open.restype = C.c_void_p
    lib.emul_stats_open.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
    lib.emul_stats_batch.restype = C.c_longlong
    lib.emul_stats_batch.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong]
    lib.emul_stats_totals.argtypes = [C.c_void_p] * 5
    lib.emul_stats_close.argtypes = [C.c_void_p]
    lib.emul_close.argtypes = [C.c_void_p]
    err = C.create_string_buffer(256)
    h = C.c_void_p(lib.emul_open(small_index.prefix.encode(), err, 256))
    assert h, err.value
    gopt = _abi.GapOpt(); fx.host_lib().fqb_gap_opt_default(C.byref(gopt)); gopt.trim_qual = 15
    cwd = os.getcwd()
    os.chdir(small_index.dir)                     # the index's .param names its side files relative to its directory
    try:
        st = C.c_void_p(lib.emul_stats_open(h, small_index.prefix.encode(), C.byref(gopt), err, 256))
        assert st, err.value
    finally:
        os.chdir(cwd)
    try:
        text = b""
        n_add = 0
        for b in range(n // batch):
            rows = np.zeros(2 * batch, _abi.READ_DTYPE)
            for e in (0, 1):
                rows[e::2] = g["b%d_e%d_rows3" % (b, e)]
            add = np.zeros(2 * batch, np.uint8)
            buf = C.create_string_buffer(batch * 256)
            k = lib.emul_stats_batch(st, batch, b * batch, 0, rows.ctypes.data_as(C.c_void_p), add.ctypes.data_as(C.c_void_p), buf, len(buf))
            assert k >= 0
            text += buf.raw[:k]
            n_add += int(add.sum())
        d = os.path.join(GOLD, "stats_" + name)
        want = open(os.path.join(d, "InsertSizeTable"), "rb").read()
        assert text.splitlines() == want.splitlines()
        assert len(want.splitlines()) > 100 and n_add > 100
        isize = np.zeros(4096, np.uint64); fsc = np.zeros(5, np.uint64); scal = np.zeros(16, np.uint64); dup = C.c_ulonglong()
        lib.emul_stats_totals(st, isize.ctypes.data_as(C.c_void_p), fsc.ctypes.data_as(C.c_void_p), scal.ctypes.data_as(C.c_void_p), C.byref(dup))
        raw = [l.split("\t") for l in open(os.path.join(d, "RawInsertSizeDist")).read().splitlines()]
        assert [int(r[0]) for r in raw] == list(range(4096))
        assert [int(r[1]) for r in raw] == [int(v) for v in isize]
        assert int(fsc[4]) == 2 * n * arrs[0].shape[1]
    finally:
        lib.emul_stats_close(st)
        lib.emul_close(h)


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_emulated_statistics_files_against_golden(name, tmp_path):
    """Rows a12-a14 end to end without a GPU: classify_pair (device function) decides which reads are added, a serial
    restatement of bases_kernel's per-base walk fills the accumulators over the side tables of build_stats_tables, and the
    product's write_summary_files writes the 12 files - compared with the reference's own (integers exact, floats 1e-9)."""
    from test_golden import _same_text
    small_index = make_golden.index_for(name)
    arrs, g, n, batch = _case(small_index, name)
    lib = fx.build_emul()
    lib.emul_open.restype = C.c_void_p
    lib.emul_stats_open.restype = C.c_void_p
    lib.emul_stats_open.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
    lib.emul_stats_batch.restype = C.c_longlong
    lib.emul_stats_batch.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong]
    lib.emul_stats_bases.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emul_stats_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.emul_stats_close.argtypes = [C.c_void_p]
    lib.emul_close.argtypes = [C.c_void_p]
    err = C.create_string_buffer(256)
    h = C.c_void_p(lib.emul_open(small_index.prefix.encode(), err, 256))
    assert h, err.value
    gopt = _abi.GapOpt(); fx.host_lib().fqb_gap_opt_default(C.byref(gopt)); gopt.trim_qual = 15
    mine = str(tmp_path / "mine")
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    st = None
    try:
        st = C.c_void_p(lib.emul_stats_open(h, small_index.prefix.encode(), C.byref(gopt), err, 256))
        assert st, err.value
        rl = arrs[0].shape[1]
        with open(mine + ".InsertSizeTable", "wb") as table:
            for b in range(n // batch):
                rows = np.zeros(2 * batch, _abi.READ_DTYPE)
                for e in (0, 1):
                    rows[e::2] = g["b%d_e%d_rows3" % (b, e)]
                add = np.zeros(2 * batch, np.uint8)
                buf = C.create_string_buffer(batch * 256)
                k = lib.emul_stats_batch(st, batch, b * batch, 1, rows.ctypes.data_as(C.c_void_p), add.ctypes.data_as(C.c_void_p), buf, len(buf))
                assert k >= 0
                table.write(buf.raw[:k])
                sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
                codes = np.zeros((2 * batch, rl), np.uint8); quals = np.zeros((2 * batch, rl), np.uint8)
                codes[0::2] = fx.NT4[sub[0]]; codes[1::2] = fx.NT4[sub[2]]
                quals[0::2] = sub[1]; quals[1::2] = sub[3]
                assert lib.emul_stats_bases(st, h, 2 * batch, rl, codes.ctypes.data_as(C.c_void_p), quals.ctypes.data_as(C.c_void_p),
                                            rows.ctypes.data_as(C.c_void_p), add.ctypes.data_as(C.c_void_p)) == 0
        assert lib.emul_stats_finish(st, C.byref(gopt), mine.encode(), b"r1.fq", b"r2.fq", err, 256) == 0, err.value
    finally:
        os.chdir(cwd)
        if st: lib.emul_stats_close(st)
        lib.emul_close(h)
    d = os.path.join(GOLD, "stats_" + name)
    for ext in make_golden.STAT_FILES:
        _same_text(os.path.join(d, ext), mine + "." + ext, sort_lines=(ext == "SexChromInfo"))
