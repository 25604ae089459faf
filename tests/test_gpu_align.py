"""GPU parity: prep + k-mer filter + bwt_cal_width + bwt_match_gap through the C ABI
against the reference's own code (oracle/_ref) on the same seeded reads.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

pytestmark = pytest.mark.gpu
CAP = 8


def _engine(index, trim_qual=15, kmer_thresh=3):
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual, g.kmer_thresh = trim_qual, kmer_thresh
    h = C.c_void_p()
    rc = lib.fqb_create(index.prefix.encode(), C.byref(g), None, 0, C.byref(h))
    assert rc == 0, lib.fqb_last_error()
    return lib, h


def _run_cuda(lib, h, arrs):
    n, L = arrs[0].shape
    rc = lib.fqb_stage_load(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None, 0)
    assert rc == 0, lib.fqb_last_error()
    rc = lib.fqb_stage_align(h)
    assert rc == 0, lib.fqb_last_error()
    ln = np.zeros(2 * n, np.int32); fl = np.zeros(2 * n, np.int32); filt = np.zeros(2 * n, np.uint8)
    codes = np.zeros((2 * n, L), np.uint8)
    assert lib.fqb_stage_fetch_prep(h, _abi.i32p(ln), _abi.i32p(fl), _abi.u8p(filt), _abi.u8p(codes), L) == 0, lib.fqb_last_error()
    aln = np.zeros((2 * n, CAP), _abi.ALN_DTYPE); na = np.zeros(2 * n, np.int32)
    assert lib.fqb_stage_fetch_aln(h, CAP, aln.ctypes.data_as(C.c_void_p), _abi.i32p(na)) == 0, lib.fqb_last_error()
    return ln, fl, filt, codes, aln, na


def _compare_with_ref(index, arrs, tag, trim_qual=15):
    fq = index.write_fastq(tag, arrs)
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=trim_qual)
    n = ref.next_batch()
    assert n == arrs[0].shape[0]
    lib, h = _engine(index, trim_qual=trim_qual)
    try:
        ln, fl, filt, codes, aln, na = _run_cuda(lib, h, arrs)
    finally:
        lib.fqb_destroy(h)
    for e in (0, 1):
        rows = ref.rows(0, e)
        sel = slice(e, None, 2)
        np.testing.assert_array_equal(ln[sel], rows["len"])
        np.testing.assert_array_equal(fl[sel], rows["full_len"])
        np.testing.assert_array_equal(filt[sel], rows["filtered"])
        rc = ref.seq_codes(e)
        L = codes.shape[1]
        for r in range(n):                      # codes agree over the trimmed length the reference kept
            k = rows["len"][r]
            assert (codes[sel][r, :k] == rc[r, :k]).all()
        off, a = ref.aln(e)
        pad, cnt = fx.csr_to_padded(off, a, CAP)
        keep = rows["filtered"] == 0
        np.testing.assert_array_equal(na[sel][keep], cnt[keep])
        assert (aln[sel][keep] == pad[keep]).all()
    return na


def test_align_on_target_2x100(small_index, ref_required):
    arrs = small_index.reads(6000, read_len=100, seed=11)
    na = _compare_with_ref(small_index, arrs, "g100")
    assert (na > 0).mean() > 0.9


def test_align_2x150_high_error(small_index, ref_required):
    arrs = small_index.reads(3000, read_len=150, seed=12, sub_rate=0.04, ins_rate=0.01, del_rate=0.01, max_indel_len=3)
    _compare_with_ref(small_index, arrs, "g150")


def test_align_mixed_offtarget_filter(small_index, ref_required):
    arrs = small_index.reads(4000, read_len=100, seed=13, f_on=0.3)
    _compare_with_ref(small_index, arrs, "gmix")


def test_align_no_trim(small_index, ref_required):
    arrs = small_index.reads(2000, read_len=100, seed=14)
    _compare_with_ref(small_index, arrs, "gnotrim", trim_qual=0)


def test_overflow_tiers_give_the_same_hits(small_index, ref_required, monkeypatch):
    """Shrunken arenas push most reads through the deeper overflow tiers; results must not change."""
    monkeypatch.setenv("FQB_DEBUG_ARENA_FAST", "48")
    monkeypatch.setenv("FQB_DEBUG_ARENA_MID", "300")
    arrs = small_index.reads(1500, read_len=100, seed=16)
    _compare_with_ref(small_index, arrs, "gover")
    lib, h = _engine(small_index)
    try:
        _run_cuda(lib, h, arrs)
        c = (C.c_uint64 * 4)()
        assert lib.fqb_stage_counters(h, c) == 0
        assert c[3] > 100          # reads that overflowed the fast pass
    finally:
        lib.fqb_destroy(h)


def test_counters_and_determinism(small_index):
    arrs = small_index.reads(2000, read_len=100, seed=15)
    lib, h = _engine(small_index)
    try:
        a = _run_cuda(lib, h, arrs)
        b = _run_cuda(lib, h, arrs)
        for x, y in zip(a, b):
            assert (x == y).all()
        c = (C.c_uint64 * 4)()
        assert lib.fqb_stage_counters(h, c) == 0
        assert c[0] > 0 and c[1] > 0
    finally:
        lib.fqb_destroy(h)


def test_memory_path_variants_give_the_same_hits(small_index, ref_required, monkeypatch):
    """FQB_SEARCH_VAR selects how the fast search pass moves its stack entries (SearchLane kVar: bit 0 pop staging through
    shared memory, bit 2 streaming stores; the build's default is 5); every form must produce the reference's hit lists.
    High-error 150-base reads keep the stacks busy (thousands of pops per read, several score buckets)."""
    arrs = small_index.reads(1500, read_len=150, seed=21, sub_rate=0.04, ins_rate=0.01, del_rate=0.01, max_indel_len=3)
    monkeypatch.setenv("FQB_SEARCH_VAR", "0")
    base = _compare_with_ref(small_index, arrs, "gvar")
    lib, h = _engine(small_index)
    try:
        monkeypatch.setenv("FQB_SEARCH_VAR", "0")
        ref = _run_cuda(lib, h, arrs)
        assert (ref[5] == base).all()
        for var in (1, 4, 5):
            monkeypatch.setenv("FQB_SEARCH_VAR", str(var))
            got = _run_cuda(lib, h, arrs)
            assert (got[5] == ref[5]).all(), var
            keep = np.arange(CAP)[None, :] < np.clip(ref[5], 0, CAP)[:, None]
            assert (got[4][keep] == ref[4][keep]).all(), var
    finally:
        lib.fqb_destroy(h)
