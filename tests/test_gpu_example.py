"""BASELINE.json configs[0]: the reference's own example (example/example.sh; ctest Test2, CMakeLists.txt:43-55) --
ERR013170 FASTQ pair (251 records per end, 151 bp and one 137-bp read, lower-case bases) vs ref.test.fa with the
hapmap.test.vcf.gz marker -- through `FASTQuick_b200 align --fq_list`, against what the reference's `align` wrote on the same
index (tests/golden/example, made by tests/golden/make_example.py).  And reads of different lengths in one batch."""
import ctypes as C
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import bamio
import fx
from fastquick_b200 import _abi
from test_gpu_stats import TEXT_FILES, _compare_files
from test_gpu_cli import CLI, _compare_bams

pytestmark = pytest.mark.gpu
EXAMPLE = os.path.join(fx.HERE, "golden", "example")


def test_reference_example_through_the_cli(tmp_path):
    work = str(tmp_path / "ex")
    shutil.copytree(EXAMPLE, work)
    cmd = [CLI, "align", "--fq_list", "fq.test.list", "--index_prefix", "test_out_ref", "--out_prefix", "b200_out"]
    r = subprocess.run(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    for msg in ("502 sequences are loaded", "500 sequences are filtered", "1 sequences are discarded of low mapQ", "1 sequences are retained for QC"):
        assert msg in r.stdout, r.stdout[-2000:]
    ref, mine = os.path.join(work, "ref_out"), os.path.join(work, "b200_out")
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(ref + "." + ext, mine + "." + ext)
    va = [l for l in open(ref + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(mine + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    recs = _compare_bams(ref + ".bam", mine + ".bam")
    assert len(recs) == 2 and recs[1]["cigar"] == "137M"


def _write_fastq(path, which, bases, quals, lens):
    with gzip.open(path, "wb", compresslevel=1) as f:
        for i in range(len(lens)):
            n = int(lens[i])
            f.write(b"@r%011d/%d\n" % (i, which))
            f.write(bases[i, :n].tobytes()); f.write(b"\n+\n"); f.write(quals[i, :n].tobytes()); f.write(b"\n")


def test_mixed_read_lengths_match_reference(small_index, ref_required):
    """Reads of 96..150 bases in one batch (lens1/lens2 of the ABI, variable-length records through the feeder): the drop-in
    must still equal the reference file by file and BAM record by BAM record."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    n = 4000
    arrs = small_index.reads(n, read_len=150, seed=313, ins_rate=0.003, del_rate=0.003)
    rng = np.random.default_rng(11)
    lens = [rng.integers(96, 151, n), rng.integers(96, 151, n)]
    lens[0][:50] = 150; lens[1][:50] = 150                     # the reference sizes its buffers by the first record
    fq = [os.path.join(small_index.dir, "mixed_%d.fq.gz" % (e + 1)) for e in (0, 1)]
    for e in (0, 1):
        _write_fastq(fq[e], e + 1, arrs[2 * e], arrs[2 * e + 1], lens[e])
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "mixed_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert len(recs) > 7000
    assert len({len(r["seq"]) for r in recs}) > 40              # many distinct read lengths made it into the BAM


def test_reads_shorter_than_the_filter_window_are_refused(small_index):
    """A read under 96 bases: the reference's k-mer filter would read stale buffer bytes; the product says so."""
    lib = fx.host_lib()
    g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    n = 200
    arrs = small_index.reads(n, read_len=100, seed=3)
    lens = [np.full(n, 100, np.int32), np.full(n, 100, np.int32)]
    lens[1][17] = 60
    rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
    rc = lib.fqb_align_pairs(h, n, 100, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), _abi.i32p(lens[0]), _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), _abi.i32p(lens[1]),
                             rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None)
    assert rc == -5 and b"shorter than 96 bases" in lib.fqb_last_error()
    # with the filter off the same batch is fine
    lib.fqb_destroy(h)
    g.kmer_thresh = 0
    assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    rc = lib.fqb_align_pairs(h, n, 100, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), _abi.i32p(lens[0]), _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), _abi.i32p(lens[1]),
                             rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None)
    assert rc == 0, lib.fqb_last_error()
    assert rows[1]["full_len"][17] == 60
    lib.fqb_destroy(h)
