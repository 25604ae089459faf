"""The pipelined form of the per-batch body (fqb_submit_pairs / fqb_collect_pairs: batch n+1's align stage on a second
stream while batch n is paired, nothing waits per batch) must give what the stage-by-stage calls give: result rows of
every batch and all statistics files.  Also: a small batch followed by a full-sized one on the same handle (the per-pair
statistics buffer must grow with the batch buffers), and fqb_prefetch_pairs in its documented call order."""
import ctypes as C
import os

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi
from test_gpu_sharded import STAT_FILES, _handle, _batch

pytestmark = pytest.mark.gpu


def _files_equal(a_prefix, b_prefix):
    for ext in STAT_FILES:
        a = [l for l in open(a_prefix + "." + ext) if not l.startswith("##fileDate")]
        b = [l for l in open(b_prefix + "." + ext) if not l.startswith("##fileDate")]
        assert a == b, ext


def test_pipelined_equals_staged(small_index, tmp_path):
    lib = fx.host_lib()
    n_batches, batch = 5, 3000
    arrs = small_index.reads(n_batches * batch, read_len=100, seed=977, ins_rate=0.004, del_rate=0.004)
    subs = [[np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs] for b in range(n_batches)]
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        one = str(tmp_path / "staged")
        h = _handle(small_index, one)
        rows_ref = [_batch(lib, h, subs[b], b * batch, None)[1] for b in range(n_batches)]
        assert lib.fqb_stats_finish(h, one.encode()) == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)

        two = str(tmp_path / "piped")
        h = _handle(small_index, two)
        rows = [[np.zeros(batch, _abi.READ_DTYPE) for _ in range(2)] for _ in range(n_batches)]

        def submit(b):
            s = subs[b]
            assert lib.fqb_submit_pairs(h, batch, 100, _abi.u8p(s[0]), _abi.u8p(s[1]), None, _abi.u8p(s[2]), _abi.u8p(s[3]), None, 0) == 0, lib.fqb_last_error()

        submit(0)
        for b in range(n_batches):
            if b + 1 < n_batches:
                submit(b + 1)
            assert lib.fqb_collect_pairs(h, rows[b][0].ctypes.data_as(C.c_void_p), rows[b][1].ctypes.data_as(C.c_void_p)) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_emit(h, None, 0) == 0, lib.fqb_last_error()      # InsertSizeTable lines of the batch collected last
        assert lib.fqb_rows_wait(h) == 0, lib.fqb_last_error()
        assert lib.fqb_stats_finish(h, two.encode()) == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)
    finally:
        os.chdir(cwd)
    for b in range(n_batches):
        for e in (0, 1):
            assert rows[b][e].tobytes() == rows_ref[b][e].tobytes(), "batch %d end %d" % (b, e)
    _files_equal(one, two)


def test_batch_buffers_grow_between_batches(small_index, tmp_path):
    """A first batch of 2,000 pairs sizes the buffers for 65,536 pairs; a later batch of 70,000 pairs must re-size every
    per-batch buffer, the per-pair statistics records included."""
    lib = fx.host_lib()
    small = small_index.reads(2000, read_len=100, seed=5)
    big = small_index.reads(70000, read_len=100, seed=6, first_pair=2000)
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        a = str(tmp_path / "grow")
        h = _handle(small_index, a)
        _batch(lib, h, small, 0, None)
        _, rows_big = _batch(lib, h, big, 2000, None)
        assert lib.fqb_stats_finish(h, a.encode()) == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)
        # the same two batches on a handle whose buffers were sized for the big batch from the start
        b = str(tmp_path / "presized")
        h = _handle(small_index, b)
        n, L = big[0].shape
        assert lib.fqb_stage_load(h, n, L, _abi.u8p(big[0]), _abi.u8p(big[1]), None, _abi.u8p(big[2]), _abi.u8p(big[3]), None, 0) == 0
        _batch(lib, h, small, 0, None)
        _, rows_big2 = _batch(lib, h, big, 2000, None)
        assert lib.fqb_stats_finish(h, b.encode()) == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)
    finally:
        os.chdir(cwd)
    for e in (0, 1):
        assert rows_big[e].tobytes() == rows_big2[e].tobytes()
    _files_equal(a, b)


def test_prefetch_is_picked_up(small_index):
    """prefetch(n+1) before align(n), as INTEGRATION.md documents it: every batch after the first two is found uploaded."""
    lib = fx.host_lib()
    lib.fqb_prefetch_hits.restype = C.c_uint64
    g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    n_batches, batch = 6, 1000
    arrs = small_index.reads(n_batches * batch, read_len=100, seed=41)
    subs = [[np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs] for b in range(n_batches)]
    rows_pre, rows_plain = [], []
    for use_prefetch in (True, False):
        assert lib.fqb_reset_stream(h) == 0
        for b in range(n_batches):
            if use_prefetch and b + 1 < n_batches:
                s = subs[b + 1]
                assert lib.fqb_prefetch_pairs(h, batch, 100, _abi.u8p(s[0]), _abi.u8p(s[1]), None, _abi.u8p(s[2]), _abi.u8p(s[3]), None) == 0, lib.fqb_last_error()
            s = subs[b]
            rows = [np.zeros(batch, _abi.READ_DTYPE) for _ in range(2)]
            assert lib.fqb_align_pairs(h, batch, 100, _abi.u8p(s[0]), _abi.u8p(s[1]), None, _abi.u8p(s[2]), _abi.u8p(s[3]), None,
                                       rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0, lib.fqb_last_error()
            (rows_pre if use_prefetch else rows_plain).append(rows)
        if use_prefetch:
            assert lib.fqb_prefetch_hits(h) >= n_batches - 2, "prefetched uploads were not picked up: %d" % lib.fqb_prefetch_hits(h)
    lib.fqb_destroy(h)
    for b in range(n_batches):
        for e in (0, 1):
            assert rows_pre[b][e].tobytes() == rows_plain[b][e].tobytes()
