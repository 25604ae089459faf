"""GPU checks of the cases added after the round's last GPU run.  On the CPU the oracle and the host-instantiated device
functions already match them (tests/test_golden.py, tests/test_emul_golden.py); this file sorts last, so whatever it finds
cannot hide the results of the other GPU tests.
  pe100_repeats: an index with planted exact repeats - REPEAT-type reads, several occurrences per SA interval, the random
                 pick among them, mapQ 0, pairing over many positions, XA lists;
  pe100_higherr: 4 % substitutions + 1 % + 1 % indels - three reads in four end unaligned and infer_isize fails in every
                 batch, so the pair stage runs with the insert-size estimate unset from the first batch on."""
import pytest

import make_golden_path  # noqa: F401  (puts tests/golden on sys.path)
import make_golden
from test_golden import _cuda_case


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.GPU_LATE))
def test_cuda_path_against_late_golden(name, tmp_path):
    _cuda_case(name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("single_end", [False, True])
def test_cli_on_repeat_index_matches_reference(ref_required, single_end):
    """`FASTQuick_b200 align` against `FASTQuick_ref align` on the repeat index, paired and single-end: every summary file,
    and every BAM record incl. XT:A:R, X0 > 1 and the XA lists of the other occurrences (about one record in ten)."""
    import os
    import subprocess

    import fx
    from test_gpu_cli import CLI, _compare_bams
    from test_gpu_stats import TEXT_FILES, _compare_files
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    index = make_golden.index_for("pe100_repeats")
    arrs = index.reads(3000, read_len=100, seed=91, f_on=0.95)
    fq = index.write_fastq("clirep", arrs)
    idx_prefix = index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(index.dir, "clirep_%s_%d" % (tag, single_end))
        cmd = [exe, "align", "--fastq_1", fq[0]] + ([] if single_end else ["--fastq_2", fq[1]]) + \
              ["--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert sum(1 for r in recs if "XA" in r["tags"]) > 200
    assert sum(1 for r in recs if r["tags"].get("XT") == ("A", "R")) > 200
    for ext in TEXT_FILES:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
