"""Shared fixtures/helpers for the test-suite (not collected as tests).

Three independent implementations meet here:
  * ref   -- the reference's own code, compiled into oracle/_ref/libfqref.so (oracle/Makefile.ref)
  * orc   -- the plain-C restatement oracle/fq_oracle.c
  * cuda  -- the product path through the C ABI (include/fastquick_b200.h)
plus `emul`, the host instantiation of the per-lane device functions (tests/emul), used only
to check kernel logic where no GPU is present.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from fastquick_b200 import _abi  # noqa: E402

REF_LIB = os.path.join(REPO, "oracle", "_ref", "libfqref.so")
REF_BIN = os.path.join(REPO, "oracle", "_ref", "FASTQuick_ref")
ORC_LIB = os.path.join(REPO, "oracle", "libfqoracle.so")
EMUL_LIB = os.path.join(HERE, "emul", "libfqemul.so")
CACHE = os.environ.get("FQB_TEST_CACHE", "/tmp/fqb_test_cache")


def have_ref():
    return os.path.exists(REF_LIB)


def build_oracle():
    srcs = [os.path.join(REPO, "oracle", f) for f in ("fq_oracle.c", "fq_oracle_pe.c", "fq_oracle_dp.c", "fq_oracle_regions.c")]
    deps = srcs + [os.path.join(REPO, "oracle", "fq_oracle.h"), os.path.join(REPO, "include", "fastquick_b200.h")]
    if not os.path.exists(ORC_LIB) or os.path.getmtime(ORC_LIB) < max(os.path.getmtime(s) for s in deps):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared"] + srcs + ["-o", ORC_LIB, "-lm"])
    return C.CDLL(ORC_LIB)


def build_emul():
    srcs = [os.path.join(HERE, "emul", "fq_emul.cpp")] + [
        os.path.join(REPO, "fastquick_b200", "csrc", f) for f in ("fq_index.cpp", "fq_relayout.cpp", "fq_hostmath.cpp", "fq_stats_host.cpp")]
    import glob
    deps = srcs + glob.glob(os.path.join(REPO, "fastquick_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(REPO, "fastquick_b200", "csrc", "*.h")) + \
        [os.path.join(REPO, "include", "fastquick_b200.h")]
    if not os.path.exists(EMUL_LIB) or os.path.getmtime(EMUL_LIB) < max(os.path.getmtime(s) for s in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", EMUL_LIB] + srcs + ["-lpthread"])
    return C.CDLL(EMUL_LIB)


_host_lib = None


def host_lib():
    """The product C-ABI library (host-side entry points work without a GPU)."""
    global _host_lib
    if _host_lib is None:
        _host_lib = _abi.load_library()
    return _host_lib


class SynthIndex:
    """A small synthetic reduced reference, built by the product's fixture builder
    (validated byte-for-byte against `FASTQuick_ref index` in test_index_build.py)."""

    def __init__(self, name, n_long=40, n_short=160, n_x=5, n_y=5, seed=0x5EED0001, with_rollhash=True, n_dup=0):
        lib = host_lib()
        self.dir = os.path.join(CACHE, name)
        os.makedirs(self.dir, exist_ok=True)
        cfg = _abi.SynthRefCfg()
        lib.fqb_synth_ref_cfg_default(C.byref(cfg))
        cfg.seed, cfg.n_long, cfg.n_short, cfg.n_x, cfg.n_y = seed, n_long, n_short, n_x, n_y
        cfg.n_dup = n_dup                                    # planted 200-base repeats between flanks
        self.cfg = cfg
        self.h = C.c_void_p()
        assert lib.fqb_synth_create(C.byref(cfg), C.byref(self.h)) == 0, lib.fqb_last_error()
        self.prefix = os.path.join(self.dir, "idx.FASTQuick.fa")
        stamp = os.path.join(self.dir, ".done_rh" if with_rollhash else ".done")
        if not os.path.exists(stamp):
            assert lib.fqb_synth_write_inputs(self.h, self.dir.encode()) == 0, lib.fqb_last_error()
            rc = lib.fqb_synth_write_index(self.h, os.path.join(self.dir, "genome.fa").encode(),
                                           os.path.join(self.dir, "dbsnp.vcf").encode(), self.prefix.encode(),
                                           1 if with_rollhash else 0)
            assert rc == 0, lib.fqb_last_error()
            open(stamp, "w").close()

    def reads(self, n_pairs, read_len=100, seed=0x5EED0002, first_pair=0, **kw):
        lib = host_lib()
        rc = _abi.SynthReadCfg()
        lib.fqb_synth_read_cfg_default(C.byref(rc))
        rc.seed, rc.read_len = seed, read_len
        for k, v in kw.items():
            setattr(rc, k, v)
        arrs = [np.zeros((n_pairs, read_len), np.uint8) for _ in range(4)]
        ret = lib.fqb_synth_reads(self.h, C.byref(rc), C.c_int64(first_pair), C.c_int64(n_pairs),
                                  *[_abi.u8p(a) for a in arrs], 0)
        assert ret == 0, lib.fqb_last_error()
        return arrs  # bases1, quals1, bases2, quals2 (ASCII)

    def write_fastq(self, tag, arrs, first_pair=0):
        lib = host_lib()
        paths = []
        n, L = arrs[0].shape
        for e in (0, 1):
            p = os.path.join(self.dir, f"{tag}_{e + 1}.fq.gz")
            rc = lib.fqb_write_fastq_gz(p.encode(), e + 1, C.c_int64(first_pair), C.c_int64(n), L,
                                        _abi.u8p(arrs[2 * e]), _abi.u8p(arrs[2 * e + 1]))
            assert rc == 0, lib.fqb_last_error()
            paths.append(p)
        return paths


NT4 = np.full(256, 4, np.uint8)
for _i, _c in enumerate("ACGT"):
    NT4[ord(_c)] = _i
    NT4[ord(_c.lower())] = _i
NT4[ord("-")] = 5


class RefRun:
    """Drives oracle/_ref/libfqref.so (the reference's own code) batch by batch."""

    def __init__(self, index_prefix, fq1, fq2, kmer_thresh=3, trim_qual=0, batch_cap=0, stats_prefix=None):
        self.lib = C.CDLL(REF_LIB)
        L = self.lib
        L.fqref_open.restype = C.c_void_p
        L.fqref_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        for f in ("fqref_rows", "fqref_aln_off", "fqref_aln", "fqref_seq_codes", "fqref_md", "fqref_multi"):
            getattr(L, f).restype = C.c_void_p
        L.fqref_rows.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("fqref_aln_off", "fqref_aln", "fqref_seq_codes", "fqref_md", "fqref_multi"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.fqref_next_batch.argtypes = [C.c_void_p]
        L.fqref_open_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.fqref_open_stats.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.fqref_isize.argtypes = [C.c_void_p, C.c_void_p]
        L.fqref_read_len.argtypes = [C.c_void_p]
        L.fqref_fsc.argtypes = [C.c_void_p, C.c_void_p]
        L.fqref_finish_stats.argtypes = [C.c_void_p]
        L.fqref_maxdiff.argtypes = [C.c_int, C.c_double, C.c_double]
        L.fqref_bwt_sa.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.fqref_bwt_sa.restype = C.c_uint32
        cwd = os.getcwd()
        os.chdir(os.path.dirname(index_prefix))   # REFERENCE_PATH in .param may be relative
        try:
            self.h = L.fqref_open(index_prefix.encode(), kmer_thresh, trim_qual, batch_cap)
            if stats_prefix:
                assert L.fqref_open_stats(self.h, index_prefix.encode(), stats_prefix.encode()) == 0
        finally:
            os.chdir(cwd)
        assert L.fqref_open_reads(self.h, fq1.encode(), fq2.encode()) == 0
        self.n = 0

    def next_batch(self):
        self.n = self.lib.fqref_next_batch(self.h)
        return self.n

    def rows(self, stage, end):
        p = self.lib.fqref_rows(self.h, stage, end)
        return np.frombuffer((C.c_char * (self.n * _abi.READ_DTYPE.itemsize)).from_address(p), _abi.READ_DTYPE).copy()

    def aln(self, end):
        off = np.frombuffer((C.c_char * ((self.n + 1) * 4)).from_address(self.lib.fqref_aln_off(self.h, end)), np.int32).copy()
        tot = int(off[-1])
        if tot == 0:
            return off, np.zeros(0, _abi.ALN_DTYPE)
        a = np.frombuffer((C.c_char * (tot * 16)).from_address(self.lib.fqref_aln(self.h, end)), _abi.ALN_DTYPE).copy()
        return off, a

    def seq_codes(self, end):
        L = self.lib.fqref_read_len(self.h)
        p = self.lib.fqref_seq_codes(self.h, end)
        return np.frombuffer((C.c_char * (self.n * L)).from_address(p), np.uint8).reshape(self.n, L).copy()

    def md(self, end):
        return C.string_at(self.lib.fqref_md(self.h, end)).decode().split("\n")[: self.n]

    def multi(self, end):
        p = self.lib.fqref_multi(self.h, end)
        return np.frombuffer((C.c_char * (self.n * _abi.FQB_MAX_MULTI * 8)).from_address(p), np.uint32).reshape(
            self.n, _abi.FQB_MAX_MULTI, 2).copy()

    def isize(self):
        ii = _abi.ISize()
        self.lib.fqref_isize(self.h, C.byref(ii))
        return ii


def csr_to_padded(off, aln, cap):
    """CSR hit lists -> (n, cap) array + counts, the layout the product/emul use."""
    n = len(off) - 1
    out = np.zeros((n, cap), _abi.ALN_DTYPE)
    cnt = np.diff(off).astype(np.int32)
    for r in range(n):
        m = min(cnt[r], cap)
        out[r, :m] = aln[off[r]: off[r] + m]
    return out, cnt
