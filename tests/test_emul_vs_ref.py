"""Kernel LOGIC without a GPU: the per-lane device functions of fq_device_core.cuh, instantiated on
the host by tests/emul, against the reference's own bwa_cal_sa_reg_gap output."""
import ctypes as C

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

CAP = 8


def _emul(index):
    lib = fx.build_emul()
    lib.emul_open.restype = C.c_void_p
    err = C.create_string_buffer(256)
    h = C.c_void_p(lib.emul_open(index.prefix.encode(), err, 256))
    assert h, err.value
    return lib, h


def _check(index, arrs, tag, arena_cap=4096, search_var=0):
    fq = index.write_fastq(tag, arrs)
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=15)
    n = ref.next_batch()
    lib, h = _emul(index)
    lib.emul_search_var(search_var)
    g = _abi.GapOpt()
    fx.host_lib().fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    for e in (0, 1):
        rows = ref.rows(0, e)
        codes = ref.seq_codes(e)
        off, a = ref.aln(e)
        pad, cnt = fx.csr_to_padded(off, a, CAP)
        keep = np.where(rows["filtered"] == 0)[0]
        sub = np.ascontiguousarray(codes[keep])
        lens = np.ascontiguousarray(rows["len"][keep].astype(np.int32))
        out = np.zeros((len(keep), CAP), _abi.ALN_DTYPE); na = np.zeros(len(keep), np.int32); st = np.zeros(len(keep), np.int32)
        rc = lib.emul_align(h, C.byref(g), len(keep), sub.shape[1], _abi.u8p(sub), _abi.i32p(lens), arena_cap, CAP,
                            out.ctypes.data_as(C.c_void_p), _abi.i32p(na), _abi.i32p(st), None)
        assert rc == 0
        assert (st == 1).all()                                   # every lane ran to completion (no overflow)
        np.testing.assert_array_equal(na, cnt[keep])
        assert (out == pad[keep]).all()
    lib.emul_search_var(0)
    lib.emul_close(h)
    stage = (C.c_ulonglong * 3)()
    lib.emul_stage_stats(stage)
    return list(stage)


def test_lanes_2x100(small_index, ref_required):
    _check(small_index, small_index.reads(2500, read_len=100, seed=31), "em100")


def test_lanes_2x150_indels(small_index, ref_required):
    _check(small_index, small_index.reads(1200, read_len=150, seed=32, sub_rate=0.03, ins_rate=0.01, del_rate=0.01, max_indel_len=3), "em150")


def test_lanes_pop_staging(small_index, ref_required):
    """The build's default form of the fast pass (SearchLane kVar = 5: pops served from a per-lane staging slot that is filled
    when the entry before it in the bucket's chain is popped): same hit lists as the reference, most pops from memory are
    served from the slot, and a staged entry never differs from the arena's (entries of the bump arena are written once)."""
    arrs = small_index.reads(1200, read_len=150, seed=34, sub_rate=0.03, ins_rate=0.01, del_rate=0.01, max_indel_len=3)
    plain = _check(small_index, arrs, "emst0")
    hit, miss, stale = _check(small_index, arrs, "emst5", search_var=1)
    assert plain == [0, 0, 0]
    assert stale == 0
    assert hit > 3 * miss > 0, (hit, miss)


def test_lanes_free_list_arena(small_index, ref_required):
    # arena_cap >= 65535 selects the 32-bit-head / free-list variant used by the deepest overflow tier
    _check(small_index, small_index.reads(600, read_len=100, seed=33), "emfl", arena_cap=70000)


def test_lanes_repeats(ref_required):
    """An index with planted exact repeats: SA intervals wider than one row, hits of equal score on both copies."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    index = make_golden.index_for("pe100_repeats")
    _check(index, index.reads(1500, read_len=100, seed=78, sub_rate=0.02), "emrep")


def test_occ_and_sa_against_oracle(small_index):
    import oracle_py
    lib, h = _emul(small_index)
    lib.emul_sa.restype = C.c_uint32
    orc = oracle_py.Oracle(small_index.prefix)
    rng = np.random.default_rng(9)
    n = orc.b0.seq_len
    cnt = (C.c_uint32 * 4)(); ref4 = (C.c_uint32 * 4)()
    for which, b in ((0, orc.b0), (1, orc.b1)):
        ks = list(rng.integers(0, n + 1, 400)) + [0, 1, n - 1, n, int(b.primary), int(b.primary) - 1, int(b.primary) + 1, 0xFFFFFFFF]
        for k in ks:
            lib.emul_occ4(h, which, C.c_uint32(int(k)), cnt)
            orc.lib.orc_occ4(C.byref(b), C.c_uint32(int(k)), ref4)
            assert list(cnt) == list(ref4), (which, k)
        for k in list(rng.integers(1, n + 1, 200)):
            assert lib.emul_sa(h, which, C.c_uint32(int(k))) == orc.lib.orc_sa(C.byref(b), C.c_uint32(int(k)))
    lib.emul_close(h)
