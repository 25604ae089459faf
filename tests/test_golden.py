"""Parity against committed golden vectors (tests/golden/, produced from the reference's own code by
tests/golden/make_golden.py): the C oracle on the CPU, the CUDA path through the C ABI on the GPU.
These do not need oracle/_ref at run time."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import fx
import oracle_py
from fastquick_b200 import _abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402

FIELDS = ["pos", "sa", "c1", "c2", "score", "len", "full_len", "clip_len", "type", "strand", "filtered", "extra_flag",
          "n_mm", "n_gapo", "n_gape", "mapQ", "seQ", "n_multi"]
FIELDS_FIN = FIELDS + ["n_cigar", "has_cigar", "nm"]


def _case(index, name):
    """index: the case's own index (make_golden.index_for(name)); kept as a parameter so callers hold on to it"""
    n, kw, batch = make_golden.CASES[name]
    arrs = index.reads(n, **kw)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert int(g["crc"][0]) == make_golden.inputs_crc(arrs), "synthetic read generator changed: regenerate tests/golden"
    return arrs, g, n, batch


def _cigars_equal(a, b, tag):
    for f in FIELDS_FIN:
        np.testing.assert_array_equal(a[f], b[f], err_msg="%s field %s" % (tag, f))
    has = b["has_cigar"] != 0
    for i in np.where(has)[0]:
        k = int(b["n_cigar"][i])
        assert (a["cigar"][i][:k] == b["cigar"][i][:k]).all(), (tag, i)


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_against_golden(name):
    index = make_golden.index_for(name)
    arrs, g, n, batch = _case(index, name)
    orc = oracle_py.Oracle(index.prefix)
    for b in range(n // batch):
        sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
        lens, filt, out, na = orc.align_batch(sub, trim_qual=15, kmer_thresh=3, cap=8)
        for e in (0, 1):
            r0 = g["b%d_e%d_rows0" % (b, e)]
            sel = slice(e, None, 2)
            np.testing.assert_array_equal(lens[sel], r0["len"])
            np.testing.assert_array_equal(filt[sel], r0["filtered"])
            pad, cnt = fx.csr_to_padded(g["b%d_e%d_aln_off" % (b, e)], g["b%d_e%d_aln" % (b, e)], 8)
            keep = r0["filtered"] == 0
            np.testing.assert_array_equal(na[sel][keep], cnt[keep])
            assert (out[sel][keep] == pad[keep]).all()
        full = np.zeros(2 * batch, np.int32)
        full[0::2] = g["b%d_e0_rows0" % b]["full_len"]; full[1::2] = g["b%d_e1_rows0" % b]["full_len"]
        rows, ii, _ = orc.pe_batch(lens, full, filt, out, na, cap=8)
        gf, gi = g["b%d_isize_f" % b], g["b%d_isize_i" % b]
        assert (ii.avg, ii.std, ii.ap_prior) == tuple(gf) and (ii.low, ii.high, ii.high_bayesian) == tuple(int(x) for x in gi)
        for e in (0, 1):
            for f in FIELDS:
                np.testing.assert_array_equal(rows[e::2][f], g["b%d_e%d_rows1" % (b, e)][f], err_msg="pe b%d e%d %s" % (b, e, f))
        after_sw, fin = orc.sw_and_refine(rows, ii, sub)
        for e in (0, 1):
            for f in FIELDS:
                np.testing.assert_array_equal(after_sw[e::2][f], g["b%d_e%d_rows2" % (b, e)][f], err_msg="sw b%d e%d %s" % (b, e, f))
            _cigars_equal(fin[e::2], g["b%d_e%d_rows3" % (b, e)], "refine b%d e%d" % (b, e))


def _close(a, b):
    if a == b:
        return True
    try:
        x, y = float(a), float(b)
    except ValueError:
        return False
    return abs(x - y) <= 1e-9 * max(abs(x), abs(y), 1e-300)          # north_star: derived floats within 1e-9 relative


def _same_text(p_gold, p_mine, sort_lines=False):
    ra = [l.rstrip("\n") for l in open(p_gold) if not l.startswith("##fileDate")]
    rb = [l.rstrip("\n") for l in open(p_mine) if not l.startswith("##fileDate")]
    if sort_lines:
        ra, rb = sorted(ra), sorted(rb)
    assert len(ra) == len(rb), (p_gold, len(ra), len(rb))
    for i, (x, y) in enumerate(zip(ra, rb)):
        if x == y:
            continue
        for ch in ", []/":
            x, y = x.replace(ch, "\t"), y.replace(ch, "\t")
        fa, fb = x.split("\t"), y.split("\t")
        assert len(fa) == len(fb) and all(_close(u, w) for u, w in zip(fa, fb)), (os.path.basename(p_gold), i, x, y)


def _cuda_case(name, tmp_path):
    small_index = make_golden.index_for(name, with_rollhash=make_golden.INDEX_OF.get(name) is None)
    arrs, g, n, batch = _case(small_index, name)
    lib = fx.host_lib()
    go = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(go))
    go.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(small_index.prefix.encode(), C.byref(go), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    mine = str(tmp_path / "mine")
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        assert lib.fqb_stats_open(h, small_index.prefix.encode()) == 0, lib.fqb_last_error()
        assert lib.fqb_stats_begin_file(h, mine.encode(), b"r1.fq", b"r2.fq") == 0, lib.fqb_last_error()
        L = arrs[0].shape[1]
        for b in range(n // batch):
            sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
            rows = [np.zeros(batch, _abi.READ_DTYPE) for _ in range(2)]
            ii = _abi.ISize()
            assert lib.fqb_align_pairs(h, batch, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None,
                                       rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), C.byref(ii)) == 0, lib.fqb_last_error()
            gf, gi = g["b%d_isize_f" % b], g["b%d_isize_i" % b]
            assert (ii.avg, ii.std, ii.ap_prior) == tuple(gf) and (ii.low, ii.high, ii.high_bayesian) == tuple(int(x) for x in gi)
            for e in (0, 1):
                _cigars_equal(rows[e], g["b%d_e%d_rows3" % (b, e)], "b%d e%d" % (b, e))
            # hit lists of the search stage
            na = np.zeros(2 * batch, np.int32); aln = np.zeros((2 * batch, 8), _abi.ALN_DTYPE)
            assert lib.fqb_stage_fetch_aln(h, 8, aln.ctypes.data_as(C.c_void_p), _abi.i32p(na)) == 0
            for e in (0, 1):
                pad, cnt = fx.csr_to_padded(g["b%d_e%d_aln_off" % (b, e)], g["b%d_e%d_aln" % (b, e)], 8)
                keep = g["b%d_e%d_rows0" % (b, e)]["filtered"] == 0
                np.testing.assert_array_equal(na[e::2][keep], cnt[keep])
                assert (aln[e::2][keep] == pad[keep]).all()
            assert lib.fqb_stage_stats(h) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_emit(h, None, 0) == 0, lib.fqb_last_error()
        assert lib.fqb_stats_finish(h, mine.encode()) == 0, lib.fqb_last_error()
    finally:
        os.chdir(cwd)
        lib.fqb_destroy(h)
    d = os.path.join(GOLD, "stats_" + name)
    for ext in make_golden.STAT_FILES:
        _same_text(os.path.join(d, ext), mine + "." + ext, sort_lines=(ext == "SexChromInfo"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c for c in make_golden.CASES if c not in make_golden.GPU_LATE])
def test_cuda_path_against_golden(name, tmp_path):
    _cuda_case(name, tmp_path)
