"""Host-side helpers of bench.py's command-line leg (no GPU): the plain-text FASTQ writer produces the records
fqb_write_fastq_gz writes, and the feeder reads them back at the counted rate."""
import ctypes as C
import gzip
import os
import sys

import numpy as np

import fx

sys.path.insert(0, fx.REPO)
import bench  # noqa: E402
from fastquick_b200 import _abi  # noqa: E402


def test_text_fastq_equals_the_gz_writer_and_feeds_back(tmp_path):
    lib = fx.host_lib()
    synth = bench.make_synth(lib)
    n, first = 3000, 262144 * 3 + 17                     # names carry the global pair index
    arrs = bench.gen_reads(lib, synth, first, n)
    fq = []
    for e in (0, 1):
        gz = str(tmp_path / ("a%d.fq.gz" % e))
        assert lib.fqb_write_fastq_gz(gz.encode(), e + 1, C.c_int64(first), C.c_int64(n), bench.READ_LEN,
                                      _abi.u8p(arrs[2 * e]), _abi.u8p(arrs[2 * e + 1])) == 0
        txt = str(tmp_path / ("a%d.fq" % e))
        bench.write_fastq_text(txt, e + 1, first, arrs[2 * e], arrs[2 * e + 1])
        assert gzip.open(gz).read() == open(txt, "rb").read()
        fq.append(txt)
    rate = bench.feeder_rate(lib, fq, n)
    assert rate > 0                                       # both files delivered exactly n records
    assert bench.feeder_rate(lib, fq, n + 1) == 0.0       # a wrong expectation is not reported as a rate
