"""Pins the plain-C oracle (oracle/fq_oracle.c) against the reference's own code
(oracle/_ref/libfqref.so = libbwa + src/BwtMapper.cpp compiled from /root/reference)."""
import ctypes as C

import numpy as np
import pytest

import fx
import oracle_py
from fastquick_b200 import _abi


@pytest.fixture(scope="module")
def ref_batch(small_index, ref_required):
    arrs = small_index.reads(1200, read_len=100, seed=21)
    fq = small_index.write_fastq("orc", arrs)
    ref = fx.RefRun(small_index.prefix, fq[0], fq[1], trim_qual=15)
    assert ref.next_batch() == 1200
    return arrs, ref


def test_prep_filter_and_hits_match_reference(small_index, ref_batch):
    arrs, ref = ref_batch
    orc = oracle_py.Oracle(small_index.prefix)
    lens, filt, out, na = orc.align_batch(arrs, trim_qual=15, kmer_thresh=3, cap=8)
    for e in (0, 1):
        rows = ref.rows(0, e)
        sel = slice(e, None, 2)
        np.testing.assert_array_equal(lens[sel], rows["len"])           # bwa_trim_read
        np.testing.assert_array_equal(filt[sel], rows["filtered"])      # BwtIndexer::IsReadFiltered
        off, a = ref.aln(e)
        pad, cnt = fx.csr_to_padded(off, a, 8)
        keep = rows["filtered"] == 0
        np.testing.assert_array_equal(na[sel][keep], cnt[keep])         # bwt_match_gap hit lists, in discovery order
        assert (out[sel][keep] == pad[keep]).all()
    assert (na > 0).mean() > 0.9


def test_maxdiff_table_matches_reference(ref_required):
    lib = fx.build_oracle()
    ref = C.CDLL(fx.REF_LIB)
    ref.fqref_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.orc_cal_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
    for l in range(0, 257):
        for thres in (0.02, 0.04, 0.001):
            assert lib.orc_cal_maxdiff(l, 0.02, thres) == ref.fqref_maxdiff(l, 0.02, thres), (l, thres)


def test_sa_lookup_matches_reference(small_index, ref_batch):
    _, ref = ref_batch
    orc = oracle_py.Oracle(small_index.prefix)
    rng = np.random.default_rng(5)
    n = orc.b0.seq_len
    for which, b in ((0, orc.b0), (1, orc.b1)):
        for k in list(rng.integers(1, n + 1, 300)) + [1, n, int(b.primary)]:
            got = orc.lib.orc_sa(C.byref(b), C.c_uint32(int(k)))
            assert got == ref.lib.fqref_bwt_sa(ref.h, which, C.c_uint32(int(k))), (which, k)


def test_drand48_stream_is_glibc(ref_required):
    """orc_drand48 must reproduce glibc's srand48(11)/drand48 stream that bwa_aln2seq_core consumes."""
    lib = fx.build_oracle()
    libc = C.CDLL("libc.so.6")
    libc.drand48.restype = C.c_double
    lib.orc_drand48.restype = C.c_double

    class Rng(C.Structure):
        _fields_ = [("x", C.c_uint64), ("n", C.c_uint64)]
    r = Rng()
    lib.orc_srand48(C.byref(r), C.c_long(11))
    libc.srand48(C.c_long(11))
    for _ in range(2000):
        assert lib.orc_drand48(C.byref(r)) == libc.drand48()


def test_multi_hit_positions_on_repeats(ref_required):
    """bwa_cal_pac_pos_pe's multi-hit lists (the XA candidates; src/BwtMapper.cpp:857-871) on an index with planted exact
    repeats: every read that keeps other hits lists the same positions, in the same order, as the reference."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    index = make_golden.index_for("pe100_repeats")
    n = 600
    arrs = index.reads(n, read_len=100, seed=77)
    fq = index.write_fastq("orcrep", arrs)
    ref = fx.RefRun(index.prefix, fq[0], fq[1], trim_qual=15)
    assert ref.next_batch() == n
    orc = oracle_py.Oracle(index.prefix)
    lens, filt, out, na = orc.align_batch(arrs, trim_qual=15, kmer_thresh=3, cap=8)
    full = np.zeros(2 * n, np.int32)
    full[0::2] = ref.rows(0, 0)["full_len"]; full[1::2] = ref.rows(0, 1)["full_len"]
    rows, ii, multi = orc.pe_batch(lens, full, filt, out, na, cap=8, want_multi=True)
    seen = 0
    for e in (0, 1):
        want, r1 = ref.multi(e), ref.rows(1, e)
        np.testing.assert_array_equal(rows[e::2]["n_multi"], r1["n_multi"])
        for i in np.where(r1["n_multi"] > 0)[0]:
            k = int(r1["n_multi"][i])
            assert want[i, :k, 0].tolist() == multi[2 * i + e, :k].tolist(), (e, i)
            seen += 1
    assert seen > 50
