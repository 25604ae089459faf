"""GPU parity for rows a12-a14: the statistics files written from the device accumulators must equal the files the
reference's own StatCollector writes (oracle/_ref, same reads, same batches) -- every integer exactly; the files are
compared as text, so the derived floats are identical to the printed precision (>= 1e-6 relative; tolerance 1e-9 is
asserted on the parsed values where a column is floating point)."""
import ctypes as C
import os

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

pytestmark = pytest.mark.gpu
TEXT_FILES = ["InsertSizeTable", "DepthDist", "GCDist", "EmpRepDist", "EmpCycleDist", "RawInsertSizeDist", "AdjustedInsertSizeDist",
              "SexChromInfo", "Pileup", "Sequence.csv", "Summary"]


def _close(a, b):
    if a == b:
        return True
    try:
        x, y = float(a), float(b)
    except ValueError:
        return False
    return abs(x - y) <= 1e-9 * max(abs(x), abs(y), 1e-300)


def _compare_files(p_ref, p_mine, sort_lines=False):
    ra = open(p_ref).read().splitlines()
    rb = open(p_mine).read().splitlines()
    if sort_lines:
        ra, rb = sorted(ra), sorted(rb)
    assert len(ra) == len(rb), (p_ref, len(ra), len(rb))
    for i, (x, y) in enumerate(zip(ra, rb)):
        if x == y:
            continue
        fa, fb = x.replace(",", "\t").replace(" ", "\t").replace("[", "\t").replace("]", "\t").replace("/", "\t").split("\t"), \
            y.replace(",", "\t").replace(" ", "\t").replace("[", "\t").replace("]", "\t").replace("/", "\t").split("\t")
        assert len(fa) == len(fb) and all(_close(u, w) for u, w in zip(fa, fb)), (os.path.basename(p_ref), i, x, y)


def _run_both(index, arrs, tag, batch, trim_qual=15):
    fq = index.write_fastq(tag, arrs)
    ref_prefix = os.path.join(index.dir, tag + "_ref")
    mine_prefix = os.path.join(index.dir, tag + "_mine")
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=trim_qual, batch_cap=batch, stats_prefix=ref_prefix)
    n_tot, L = arrs[0].shape
    for b in range(n_tot // batch):
        assert ref.next_batch() == batch
    cwd = os.getcwd()
    os.chdir(index.dir)
    try:
        assert ref.lib.fqref_finish_stats(ref.h) == 0
        lib = fx.host_lib()
        g = _abi.GapOpt()
        lib.fqb_gap_opt_default(C.byref(g))
        g.trim_qual = trim_qual
        h = C.c_void_p()
        assert lib.fqb_create(index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
        try:
            assert lib.fqb_stats_open(h, index.prefix.encode()) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_begin_file(h, mine_prefix.encode(), b"r1.fq", b"r2.fq") == 0, lib.fqb_last_error()
            for b in range(n_tot // batch):
                sub = [np.ascontiguousarray(a[b * batch:(b + 1) * batch]) for a in arrs]
                assert lib.fqb_align_pairs(h, batch, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None,
                                           None, None, None) == 0, lib.fqb_last_error()
                assert lib.fqb_stage_stats(h) == 0, lib.fqb_last_error()
                assert lib.fqb_stats_emit(h, None, 0) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_finish(h, mine_prefix.encode()) == 0, lib.fqb_last_error()
        finally:
            lib.fqb_destroy(h)
    finally:
        os.chdir(cwd)
    for ext in TEXT_FILES:
        _compare_files(ref_prefix + "." + ext, mine_prefix + "." + ext, sort_lines=(ext == "SexChromInfo"))
    # the genotype-likelihood VCF differs only in its fileDate header line
    va = [l for l in open(ref_prefix + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(mine_prefix + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    # SexChromInfo must also come out in the reference's (unordered_map) order
    assert open(ref_prefix + ".SexChromInfo").read() == open(mine_prefix + ".SexChromInfo").read()
    return ref_prefix, mine_prefix


def test_stats_files_2x100_two_batches(small_index, ref_required):
    arrs = small_index.reads(8000, read_len=100, seed=71)
    ref_prefix, _ = _run_both(small_index, arrs, "s100", 4000)
    assert sum(1 for _ in open(ref_prefix + ".Pileup")) > 100
    assert sum(1 for _ in open(ref_prefix + ".InsertSizeTable")) > 7000


def test_stats_files_indel_rich_offtarget_mix(small_index, ref_required):
    arrs = small_index.reads(5000, read_len=100, seed=72, f_on=0.8, sub_rate=0.02, ins_rate=0.003, del_rate=0.003, max_indel_len=3)
    _run_both(small_index, arrs, "smix", 5000)
