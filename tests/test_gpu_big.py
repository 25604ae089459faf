"""The bench configuration itself, pinned: 270,000 synthetic pairs (one full 262,144-pair batch and a partial one) of the
BASELINE configs[1] workload on the 10,197-marker index, `FASTQuick_b200 align` against the reference's own `align`
(oracle/_ref, ~35 s on the box's host cores): every statistics file and every BAM record."""
import os
import subprocess

import pytest

import fx
from test_gpu_stats import TEXT_FILES, _compare_files
from test_gpu_cli import CLI, _compare_bams

pytestmark = pytest.mark.gpu


def test_full_batch_on_the_10k_marker_index(ref_required):
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    index = fx.SynthIndex("bench10k_rh", n_long=1000, n_short=9000, n_x=100, n_y=97, with_rollhash=True)
    n = 270000
    arrs = index.reads(n, read_len=100)                     # the bench's read configuration (defaults of fqb_synth_read_cfg_t)
    fq = index.write_fastq("big", arrs)
    idx_prefix = index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(index.dir, "big_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", str(os.cpu_count() or 4), "--q", "15"]
        r = subprocess.run(cmd, cwd=index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    va = [l for l in open(outs["ref"] + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(outs["b200"] + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert len(recs) > 500000
    for f in fq + [outs["ref"] + ".bam", outs["b200"] + ".bam", index.prefix + ".rollhash"]:
        try:
            os.remove(f)
        except OSError:
            pass
    os.remove(os.path.join(index.dir, ".done_rh"))          # the 3 GiB k-mer table file was removed: rebuild the fixture next time
