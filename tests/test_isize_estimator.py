"""Row a14, host only: InsertSizeEstimator (src/InsertSizeEstimator.cpp:43-173) over the reference's own InsertSizeTable
reproduces the reference's AdjustedInsertSizeDist (tests/golden/stats_*, written by FASTQuick_ref through
tests/golden/make_golden.py).  The file holds the 2,000 densities at the stream's default six significant digits, so the
comparison is on the text; a parsed comparison at 1e-9 relative guards the same numbers against formatting changes."""
import ctypes as C
import os

import numpy as np
import pytest

import fx

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["stats_pe100", "stats_pe150_indel", "stats_pe100_higherr", "stats_pe100_repeats"])
def test_adjusted_insert_size_dist_matches_reference(tmp_path, case):
    lib = fx.host_lib()
    lib.fqb_isize_adjusted_file.argtypes = [C.c_char_p, C.c_char_p]
    out = str(tmp_path / "adj")
    assert lib.fqb_isize_adjusted_file(os.path.join(GOLD, case, "InsertSizeTable").encode(), out.encode()) == 0
    want = open(os.path.join(GOLD, case, "AdjustedInsertSizeDist")).read()
    got = open(out).read()
    assert got == want
    a = np.array([float(l.split("\t")[1]) for l in got.splitlines()])
    b = np.array([float(l.split("\t")[1]) for l in want.splitlines()])
    assert a.shape == (2000,) and np.allclose(a, b, rtol=1e-9, atol=0) and a.max() > 0


def test_adjusted_insert_size_missing_table(tmp_path):
    lib = fx.host_lib()
    lib.fqb_isize_adjusted_file.argtypes = [C.c_char_p, C.c_char_p]
    assert lib.fqb_isize_adjusted_file(str(tmp_path / "nope").encode(), str(tmp_path / "o").encode()) != 0
