"""GPU parity for rows a6-a9 (bwa_cal_pac_pos_pe): SE hit choice on the glibc drand48 stream, bwt_sa positions,
SE mapQ, infer_isize and pairing, across two consecutive batches (RNG position and last_ii carry over),
against the reference's own snapshot taken right after bwa_cal_pac_pos_pe."""
import ctypes as C

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

pytestmark = pytest.mark.gpu
FIELDS = ["pos", "sa", "c1", "c2", "score", "len", "full_len", "clip_len", "type", "strand", "filtered", "extra_flag",
          "n_mm", "n_gapo", "n_gape", "mapQ", "seQ", "n_multi"]


def _run(index, arrs, tag, batch, trim_qual=15, **ref_kw):
    fq = index.write_fastq(tag, arrs)
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=trim_qual, batch_cap=batch)
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = trim_qual
    h = C.c_void_p()
    assert lib.fqb_create(index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    try:
        n_tot, L = arrs[0].shape
        for b in range(n_tot // batch):
            assert ref.next_batch() == batch
            sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
            assert lib.fqb_stage_load(h, batch, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None, 0) == 0
            assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()
            assert lib.fqb_stage_pair(h) == 0, lib.fqb_last_error()
            rows = [np.zeros(batch, _abi.READ_DTYPE) for _ in range(2)]
            ii = _abi.ISize()
            assert lib.fqb_stage_fetch_rows(h, rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), C.byref(ii)) == 0
            rii = ref.isize()
            assert (ii.avg, ii.std, ii.ap_prior, ii.low, ii.high, ii.high_bayesian) == (rii.avg, rii.std, rii.ap_prior, rii.low, rii.high, rii.high_bayesian)
            for e in (0, 1):
                rr = ref.rows(1, e)
                for f in FIELDS:
                    np.testing.assert_array_equal(rows[e][f], rr[f], err_msg="batch %d end %d field %s" % (b, e, f))
            yield rows
    finally:
        lib.fqb_destroy(h)


def test_pair_stage_two_batches(small_index, ref_required):
    arrs = small_index.reads(6000, read_len=100, seed=51)
    out = list(_run(small_index, arrs, "p100", 3000))
    assert len(out) == 2
    assert ((out[0][0]["extra_flag"] & 2) != 0).mean() > 0.8      # most pairs end up properly paired


def test_pair_stage_high_error_150(small_index, ref_required):
    arrs = small_index.reads(3000, read_len=150, seed=52, sub_rate=0.04, ins_rate=0.01, del_rate=0.01, max_indel_len=3)
    list(_run(small_index, arrs, "p150", 3000))


def test_pair_stage_mostly_filtered_batch_falls_back_to_last_isize(small_index, ref_required):
    # second batch has too few confident pairs to infer an insert size: ii must fall back to the first batch's
    a = small_index.reads(3000, read_len=100, seed=53)
    b = small_index.reads(3000, read_len=100, seed=54, f_on=0.004)
    arrs = [np.concatenate([x, y]) for x, y in zip(a, b)]
    list(_run(small_index, arrs, "pfall", 3000))


# ---- rows a10 + a11: bwa_paired_sw and bwa_refine_gapped (snapshots 2 and 3 of the reference) ----
FIELDS_FIN = FIELDS + ["n_cigar", "has_cigar", "nm"]


def _run_full(index, arrs, tag, batch, trim_qual=15):
    fq = index.write_fastq(tag, arrs)
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=trim_qual, batch_cap=batch)
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = trim_qual
    h = C.c_void_p()
    assert lib.fqb_create(index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    stats = {"matesw": 0, "cigars": 0}
    try:
        n_tot, L = arrs[0].shape
        for b in range(n_tot // batch):
            assert ref.next_batch() == batch
            sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
            assert lib.fqb_stage_load(h, batch, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None, 0) == 0
            assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()
            assert lib.fqb_stage_pair(h) == 0, lib.fqb_last_error()
            assert lib.fqb_stage_sw_refine(h) == 0, lib.fqb_last_error()
            rows = [np.zeros(batch, _abi.READ_DTYPE) for _ in range(2)]
            assert lib.fqb_stage_fetch_rows(h, rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0
            for e in (0, 1):
                rr = ref.rows(3, e)
                for f in FIELDS_FIN:
                    np.testing.assert_array_equal(rows[e][f], rr[f], err_msg="batch %d end %d field %s" % (b, e, f))
                for r in np.where(rr["has_cigar"] != 0)[0]:
                    k = rr["n_cigar"][r]
                    assert list(rows[e]["cigar"][r][:k]) == list(rr["cigar"][r][:k]), (b, e, r)
                stats["matesw"] += int((rr["type"] == 3).sum())
                stats["cigars"] += int((rr["has_cigar"] != 0).sum())
    finally:
        lib.fqb_destroy(h)
    return stats


def test_sw_and_refine_2x100(small_index, ref_required):
    st = _run_full(small_index, small_index.reads(6000, read_len=100, seed=61), "f100", 3000)
    assert st["matesw"] > 100 and st["cigars"] > 1000


def test_sw_and_refine_indel_rich_150(small_index, ref_required):
    arrs = small_index.reads(3000, read_len=150, seed=62, sub_rate=0.02, ins_rate=0.004, del_rate=0.004, max_indel_len=3)
    st = _run_full(small_index, arrs, "f150", 3000)
    assert st["matesw"] > 200


def test_sw_long_indels_near_read_ends(small_index, ref_required):
    """Mate rescue of reads whose indel sits a few bases from the read start: the forward pass of aln_local_core cuts
    the gap chain there (h < q + r), the reverse pass does not, so the reverse score overtakes score_f + q + r and the
    reference's stop rule must be reproduced as a running maximum (regression test for the warp-cooperative kernel)."""
    arrs = small_index.reads(6000, read_len=100, seed=84, sub_rate=0.03, ins_rate=0.006, del_rate=0.006, max_indel_len=3)
    st = _run_full(small_index, arrs, "fends", 3000)
    assert st["matesw"] > 300


def test_warp_and_per_lane_dp_kernels_agree(small_index):
    """The one-alignment-per-warp kernels and the per-lane kernels (FQB_DP_NO_WARP, the retry path) give identical rows."""
    import os
    import zlib
    arrs = small_index.reads(8000, read_len=100, seed=85, sub_rate=0.03, ins_rate=0.008, del_rate=0.008, max_indel_len=4)
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    crcs = []
    for no_warp in (False, True):
        if no_warp:
            os.environ["FQB_DP_NO_WARP"] = "1"
        try:
            h = C.c_void_p()
            assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
            n, L = arrs[0].shape
            rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
            assert lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None,
                                       rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0, lib.fqb_last_error()
            lib.fqb_destroy(h)
            crcs.append((zlib.crc32(rows[0].tobytes()), zlib.crc32(rows[1].tobytes())))
        finally:
            os.environ.pop("FQB_DP_NO_WARP", None)
    assert crcs[0] == crcs[1]
    assert int((rows[0]["type"] == 3).sum() + (rows[1]["type"] == 3).sum()) > 500


def test_async_row_fetch_equals_sync(small_index):
    arrs = small_index.reads(3000, read_len=100, seed=86)
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    try:
        n, L = arrs[0].shape
        sync = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
        assert lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None,
                                   sync[0].ctypes.data_as(C.c_void_p), sync[1].ctypes.data_as(C.c_void_p), None) == 0, lib.fqb_last_error()
        asy = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
        assert lib.fqb_stage_fetch_rows_async(h, asy[0].ctypes.data_as(C.c_void_p), asy[1].ctypes.data_as(C.c_void_p)) == 0, lib.fqb_last_error()
        # the next batch may start right away; its pair stage overwrites the device rows only after they were split off
        assert lib.fqb_stage_load(h, n, L, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, 0) == 0
        assert lib.fqb_stage_align(h) == 0 and lib.fqb_stage_pair(h) == 0, lib.fqb_last_error()
        assert lib.fqb_rows_wait(h) == 0
        for e in (0, 1):
            assert asy[e].tobytes() == sync[e].tobytes()
    finally:
        lib.fqb_destroy(h)
