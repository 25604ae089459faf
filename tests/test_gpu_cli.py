"""Drop-in check at the CLI level: `FASTQuick_b200 align` (C++ host over the C ABI) vs the reference's own
`FASTQuick align` (oracle/_ref/FASTQuick_ref), same flags as bin/FASTQuick.sh --steps Align passes, same FASTQ files;
every summary file must be identical (BAM emission is the next row and is not compared)."""
import os
import subprocess

import pytest

import fx
from test_gpu_stats import TEXT_FILES, _compare_files

pytestmark = pytest.mark.gpu
CLI = os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200")


def test_cli_align_matches_reference(small_index, ref_required):
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    arrs = small_index.reads(5000, read_len=100, seed=81)
    fq = small_index.write_fastq("cli", arrs)
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "cli_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        assert "Processed Pair End mapping in" in r.stdout
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    va = [l for l in open(outs["ref"] + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(outs["b200"] + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
