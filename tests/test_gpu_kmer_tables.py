"""Row f4: the k-mer pre-filter tables built on the device at fqb_create are, bit for bit, the 3 GiB <prefix>.rollhash that
BwtIndexer::BuildIndex writes (tests/test_index_build.py pins the fixture builder's file to the reference's own `index` run,
with and without ambiguous bases in the flanks)."""
import ctypes as C
import os

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

pytestmark = pytest.mark.gpu
TOTAL = 6 << 29


def _create(prefix, env=None):
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    h = C.c_void_p()
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        assert lib.fqb_create(prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    return lib, h


def _assert_tables_equal_file(lib, h, prefix):
    lib.fqb_kmer_tables_fetch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    want = np.memmap(prefix + ".rollhash", dtype=np.uint8, mode="r")
    assert want.size == TOTAL
    chunk = 1 << 28
    got = np.empty(chunk, np.uint8)
    n_set = 0
    for off in range(0, TOTAL, chunk):
        assert lib.fqb_kmer_tables_fetch(h, off, chunk, got.ctypes.data_as(C.c_void_p)) == 0, lib.fqb_last_error()
        w = np.asarray(want[off:off + chunk])
        if not np.array_equal(got, w):
            bad = np.flatnonzero(got != w)
            raise AssertionError("tables differ at %d bytes from offset %d, first %d: %02x vs %02x" % (bad.size, off, bad[0], got[bad[0]], w[bad[0]]))
        n_set += int(np.count_nonzero(w))
    assert n_set > 0


def test_device_built_tables_equal_rollhash_file(small_index):
    lib, h = _create(small_index.prefix)
    try:
        assert lib.fqb_kmer_tables_origin(h) == 1            # built on the device, the file was not read
        _assert_tables_equal_file(lib, h, small_index.prefix)
    finally:
        lib.fqb_destroy(h)
    lib, h = _create(small_index.prefix, {"FQB_ROLLHASH_FROM_FILE": "1"})
    try:
        assert lib.fqb_kmer_tables_origin(h) == 3
    finally:
        lib.fqb_destroy(h)


def test_device_built_tables_with_ambiguous_bases(small_index):
    """N, IUPAC codes, '-' and lower case in the flanks: the reference substitutes rand() % 4 per visit; the draws are reproduced."""
    d = os.path.join(fx.CACHE, "kmer_amb")
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "amb.FASTQuick.fa")
    lib = fx.host_lib()
    if not os.path.exists(os.path.join(d, ".done")):
        lines = open(small_index.prefix).read().split("\n")
        rng = np.random.default_rng(5)
        n_flanks = len([l for l in lines if l.startswith(">")])
        for f in rng.choice(n_flanks, 9, replace=False):
            s = bytearray(lines[2 * f + 1].encode())
            half = len(s) // 2
            for pos, ch in ((3, b"N"), (40, b"n"), (half - 33, b"R"), (half - 1, b"N"), (half, b"N"), (half + 5, b"N"), (half + 6, b"N"),
                            (half + 31, b"Y"), (half + 32, b"N"), (len(s) - 2, b"N"), (100, b"-"))[:int(rng.integers(3, 12))]:
                s[pos] = ch[0]
            s[60:70] = bytes(s[60:70]).lower()
            lines[2 * f + 1] = s.decode()
        fa = os.path.join(d, "flanks.fa")
        open(fa, "w").write("\n".join(lines))
        assert lib.fqb_index_from_flank_fasta(fa.encode(), prefix.encode(), 1) == 0, lib.fqb_last_error()
        open(os.path.join(d, ".done"), "w").close()
    assert int(open(prefix + ".amb").readline().split()[2]) > 0
    lib, h = _create(prefix)
    try:
        assert lib.fqb_kmer_tables_origin(h) == 1
        _assert_tables_equal_file(lib, h, prefix)
    finally:
        lib.fqb_destroy(h)
