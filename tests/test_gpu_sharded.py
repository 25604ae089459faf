"""Row (e) on the real engine: two handles (as two ranks would hold them) process alternate batches with the drand48
position and last_ii handed along, their accumulators are exported, summed and imported into the first one -- the result
rows and all statistics files must equal what one handle produces over the same batches in order.
Also the full-size property checks (one reference-sized batch on the 10k-marker index)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

pytestmark = pytest.mark.gpu
STAT_FILES = ["InsertSizeTable", "DepthDist", "GCDist", "EmpRepDist", "EmpCycleDist", "RawInsertSizeDist", "AdjustedInsertSizeDist",
              "SexChromInfo", "Pileup", "Sequence.csv", "Summary", "vcf"]


def _handle(index, out_prefix):
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_open(h, index.prefix.encode()) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_begin_file(h, out_prefix.encode(), b"r1.fq", b"r2.fq") == 0, lib.fqb_last_error()
    return h


def _batch(lib, h, sub, first_pair, state_in):
    n, L = sub[0].shape
    assert lib.fqb_set_pair_base(h, C.c_uint64(first_pair)) == 0
    assert lib.fqb_stage_load(h, n, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None, 0) == 0
    assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()
    if state_in is not None:
        assert lib.fqb_set_stream_state(h, C.c_uint64(state_in[0]), C.byref(state_in[1])) == 0
    assert lib.fqb_stage_pair(h) == 0, lib.fqb_last_error()
    calls, ii = C.c_uint64(0), _abi.ISize()
    assert lib.fqb_get_stream_state(h, C.byref(calls), C.byref(ii)) == 0
    assert lib.fqb_stage_sw_refine(h) == 0, lib.fqb_last_error()
    rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
    assert lib.fqb_stage_fetch_rows(h, rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0
    assert lib.fqb_stage_stats(h) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_emit(h, None, 0) == 0, lib.fqb_last_error()
    return (calls.value, ii), rows


def test_two_handles_equal_one(small_index, tmp_path):
    import torch
    lib = fx.host_lib()
    n_batches, batch = 4, 1500
    arrs = small_index.reads(n_batches * batch, read_len=100, seed=131)
    subs = [[np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs] for b in range(n_batches)]
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        one = str(tmp_path / "one")
        h = _handle(small_index, one)
        rows_one = []
        for b in range(n_batches):
            _, rows = _batch(lib, h, subs[b], b * batch, None)
            rows_one.append(rows)
        assert lib.fqb_stats_finish(h, one.encode()) == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)

        two = str(tmp_path / "two")
        hs = [_handle(small_index, two), _handle(small_index, two + "_r1")]
        state = None
        for b in range(n_batches):
            state, rows = _batch(lib, hs[b % 2], subs[b], b * batch, state)
            for e in (0, 1):
                assert rows[e].tobytes() == rows_one[b][e].tobytes(), "batch %d end %d rows differ between the sharded and the single run" % (b, e)
        # reduce the accumulator groups of "rank 1" into "rank 0" (what NCCL does across GPUs): sum, first-touch order by min
        for which, dt, op in ((0, torch.int32, "sum"), (1, torch.int64, "sum"), (2, torch.int32, "sum"), (3, torch.int32, "min")):
            nb = C.c_uint64(0)
            assert lib.fqb_stats_group_bytes(hs[0], which, C.byref(nb)) == 0
            t = [torch.empty(int(nb.value) // (4 if dt == torch.int32 else 8), dtype=dt, device="cuda") for _ in range(2)]
            for r in range(2):
                assert lib.fqb_stats_export(hs[r], which, C.c_void_p(t[r].data_ptr())) == 0, lib.fqb_last_error()
            torch.cuda.synchronize()
            merged = t[0] + t[1] if op == "sum" else torch.minimum(t[0], t[1])
            torch.cuda.synchronize()
            assert lib.fqb_stats_import(hs[0], which, C.c_void_p(merged.data_ptr())) == 0, lib.fqb_last_error()
        # variable-size state: marker pile-up entries and the distinct PCR-duplicate keys of rank 1
        for which, item in ((0, 20), (1, 8)):
            cnt = C.c_uint64(0)
            assert lib.fqb_stats_var_count(hs[1], which, C.byref(cnt)) == 0, lib.fqb_last_error()
            buf = np.zeros(max(int(cnt.value), 1) * item, np.uint8)
            assert lib.fqb_stats_var_export(hs[1], which, buf.ctypes.data_as(C.c_void_p), C.c_uint64(cnt.value)) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_var_import(hs[0], which, buf.ctypes.data_as(C.c_void_p), C.c_uint64(cnt.value)) == 0, lib.fqb_last_error()
        # the InsertSizeTable lines of rank 1's batches are spliced back into file order on rank 0
        assert lib.fqb_stats_close_table(hs[1]) == 0, lib.fqb_last_error()
        others = (C.c_char_p * 1)((two + "_r1").encode())
        assert lib.fqb_stats_merge_tables(hs[0], others, 1) == 0, lib.fqb_last_error()
        assert lib.fqb_stats_finish(hs[0], two.encode()) == 0, lib.fqb_last_error()
        for x in hs:
            lib.fqb_destroy(x)
    finally:
        os.chdir(cwd)
    for ext in STAT_FILES:
        a = [l for l in open(one + "." + ext) if not l.startswith("##fileDate")]
        b = [l for l in open(two + "." + ext) if not l.startswith("##fileDate")]
        assert a == b, ext


def test_full_size_batch_properties():
    """One reference-sized batch (262,144 pairs, BASELINE configs[1] shape) on the 10k-marker index: every reported
    alignment is consistent with the packed reference (NM recomputed from pos/strand/CIGAR), the run is deterministic,
    and the batch-level counts are sane."""
    lib = fx.host_lib()
    cfg = _abi.SynthRefCfg(); lib.fqb_synth_ref_cfg_default(C.byref(cfg))
    s = C.c_void_p(); assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0
    g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create_from_synth(s, C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    n, L = _abi.FQB_BATCH_PAIRS, 100
    rc_ = _abi.SynthReadCfg(); lib.fqb_synth_read_cfg_default(C.byref(rc_)); rc_.read_len = L
    arrs = [np.zeros((n, L), np.uint8) for _ in range(4)]
    assert lib.fqb_synth_reads(s, C.byref(rc_), C.c_int64(0), C.c_int64(n), *[_abi.u8p(x) for x in arrs], 0) == 0
    crcs = []
    for _ in range(2):
        assert lib.fqb_reset_stream(h) == 0
        rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
        assert lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None,
                                   rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0, lib.fqb_last_error()
        crcs.append((zlib.crc32(rows[0].tobytes()), zlib.crc32(rows[1].tobytes())))
    assert crcs[0] == crcs[1], "the same batch from the same stream state must give identical rows"
    l_pac = C.c_int64(0); nc = C.c_int32(0)
    assert lib.fqb_index_info(h, C.byref(l_pac), C.byref(nc), None, None) == 0
    lib.fqb_destroy(h)
    mapped = [(r["type"] != 0) for r in rows]
    assert mapped[0].mean() > 0.95 and mapped[1].mean() > 0.95
    assert ((rows[0]["extra_flag"] & 2) != 0).mean() > 0.9                       # properly paired
    for e in (0, 1):
        r = rows[e][mapped[e]]
        assert (r["pos"].astype(np.int64) + 1 <= l_pac.value).all()
        assert (r["nm"] >= r["n_mm"].astype(np.int64) * 0).all()
        gapped = r["has_cigar"] != 0
        # CIGAR query length == full read length for every row that carries one
        ql = np.zeros(len(r), np.int64)
        for k in range(_abi.FQB_MAX_CIGAR):
            c = r["cigar"][:, k].astype(np.int64)
            use = (k < r["n_cigar"]) & gapped & ((c >> 14) != 2)
            ql += np.where(use, c & 0x3fff, 0)
        assert (ql[gapped] == r["full_len"][gapped]).all()
        assert (r["mapQ"] <= 60).all()
        assert (r["n_gapo"][r["type"] != 3] <= 1).all()       # -o 1 binds bwt_match_gap; mate-rescue (BWA_TYPE_MATESW) alignments are free
