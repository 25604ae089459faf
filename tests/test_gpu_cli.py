"""Drop-in check at the CLI level: `FASTQuick_b200 align` (C++ host over the C ABI) vs the reference's own
`FASTQuick align` (oracle/_ref/FASTQuick_ref), same flags as bin/FASTQuick.sh --steps Align passes, same FASTQ files;
every summary file must be identical, and the BAM files must hold the same header and the same records (every fixed
field and every tag; tag order inside a record is libStatGen's hash order and is not compared)."""
import os
import subprocess

import pytest

import numpy as np

import bamio
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


def _compare_bams(p_ref, p_mine):
    ta, ra, a = bamio.read_bam(p_ref)
    tb, rb, b = bamio.read_bam(p_mine)
    assert ta == tb, "BAM header text differs"
    assert ra == rb
    assert len(a) == len(b), (len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            j = i ^ 1
            strip = lambda r: {k: v for k, v in r.items() if k not in ("qual", "seq")}
            raise AssertionError((i, {k: (x[k], y[k]) for k in x if x[k] != y[k]}, strip(x), strip(y), strip(a[j]), strip(b[j])))
    return a


def test_cli_bam_matches_reference(small_index, ref_required):
    """Row f1: the records of BwtMapper::SetSamRecord.  The input mixes well-behaved pairs, indel-rich pairs (gapped
    CIGARs, MD with deletions, XA of gapped alternative hits), quality-trimmed reads (XC, soft clips), half-mapped pairs
    (one end replaced by noise: mate-unmapped flags, the unmapped read placed at its mate) and off-target pairs."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    a = small_index.reads(3000, read_len=100, seed=83, f_on=0.9)
    b = small_index.reads(1500, read_len=100, seed=84, sub_rate=0.03, ins_rate=0.006, del_rate=0.006, max_indel_len=3)
    arrs = [np.concatenate([x, y]) for x, y in zip(a, b)]
    rng = np.random.default_rng(7)
    junk = rng.choice(len(arrs[0]), 300, replace=False)
    for i in junk:                                   # one end becomes random sequence
        e = 0 if i % 2 else 2
        arrs[e][i] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, arrs[e].shape[1])]
    fq = small_index.write_fastq("clibam", arrs)
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "clibam_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert len(recs) > 7000
    assert sum(1 for r in recs if r["flag"] & 4) > 50          # unmapped reads of half-mapped pairs
    assert sum(1 for r in recs if "XC" in r["tags"]) > 300     # quality-trimmed reads
    assert sum(1 for r in recs if "D" in r["cigar"] or "I" in r["cigar"]) > 300
    for ext in TEXT_FILES:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)


def test_cli_single_end_matches_reference(small_index, ref_required):
    """Row f3: SingleEndMapper (src/BwtMapper.cpp:1266-1407) -- bwa_aln2seq_core with a multi list of up to 3 hits,
    bwa_cal_pac_pos, gapped refinement, AddAlignment(p, 0), SetSamRecord(p, 0): every summary file and every BAM record."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    a = small_index.reads(4000, read_len=100, seed=87, f_on=0.9)
    b = small_index.reads(1500, read_len=100, seed=88, sub_rate=0.03, ins_rate=0.006, del_rate=0.006, max_indel_len=3)
    arrs = [np.concatenate([x, y]) for x, y in zip(a, b)]
    fq = small_index.write_fastq("clise", arrs)
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "clise_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        assert "Processed Single End mapping in" in r.stdout
        outs[tag] = out
    for ext in TEXT_FILES:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    va = [l for l in open(outs["ref"] + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(outs["b200"] + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert len(recs) > 4000 and sum(1 for r in recs if "XC" in r["tags"]) > 300


def test_cli_frac_samp_matches_reference(small_index, ref_required):
    """--frac_samp: both implementations drop the same records (Mersenne twister re-seeded per IO round)."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    arrs = small_index.reads(6000, read_len=100, seed=89)
    fq = small_index.write_fastq("clifrac", arrs)
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "clifrac_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15",
               "--frac_samp", "0.4"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert 3500 < len(recs) < 6000          # about 40 % of 6000 pairs, two records each


def test_cli_target_region_matches_reference(small_index, ref_required, tmp_path):
    """TARGET_REGION_PATH of the index's .param (set by `index --regionList`): StatCollector::SetTargetRegion restricts the
    regular-site statistics to flankRegion (inner join) targetRegion and changes the size terms of DepthDist / Summary."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    # a copy of the index whose .param names a BED file: windows around every third marker, some overlapping, some off-flank
    d = str(tmp_path)
    base = os.path.basename(small_index.prefix)
    for f in os.listdir(small_index.dir):
        if f.startswith(base) and not f.endswith(".param"):
            os.symlink(os.path.join(small_index.dir, f), os.path.join(d, f))
    bed = os.path.join(d, "target.bed")
    with open(bed, "w") as fo:
        for i, line in enumerate(l for l in open(small_index.prefix + ".SelectedSite.vcf") if not l.startswith("#")):
            chrom, pos = line.split("\t")[:2]
            pos = int(pos)
            if i % 3 == 0:
                fo.write("%s\t%d\t%d\n" % (chrom, pos - 180, pos + 60))
                fo.write("chr%s\t%d\t%d\n" % (chrom, pos + 40, pos + 120))
            elif i % 3 == 1:
                fo.write("%s\t%d\t%d\n" % (chrom, pos + 5000, pos + 5200))
    with open(os.path.join(d, base + ".param"), "w") as fo:
        for l in open(small_index.prefix + ".param"):
            fo.write("TARGET_REGION_PATH\t%s\n" % bed if l.startswith("TARGET_REGION_PATH") else l)
    arrs = small_index.reads(5000, read_len=100, seed=90)
    fq = small_index.write_fastq("clitr", arrs)
    idx_prefix = os.path.join(d, base)[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(d, "clitr_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        assert "Read in target region from" in r.stdout
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    va = [l for l in open(outs["ref"] + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(outs["b200"] + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    # and the restriction really changed something
    full = os.path.join(small_index.dir, "cli_ref.DepthDist")
    if os.path.exists(full):
        assert open(full).read() != open(outs["ref"] + ".DepthDist").read()


def test_cli_fastq_list_matches_reference(small_index, ref_required):
    """--fq_list with two paired files and one single-end file: per-file counters (FASTQ.csv), the RNG restart per file,
    and pile-up / InsertSizeTable order across files."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    fqa = small_index.write_fastq("clila", small_index.reads(2500, read_len=100, seed=92))
    fqb = small_index.write_fastq("clilb", small_index.reads(1800, read_len=100, seed=93, f_on=0.8, sub_rate=0.02))
    fqc = small_index.write_fastq("clilc", small_index.reads(1200, read_len=100, seed=94))
    lst = os.path.join(small_index.dir, "clil.list")
    with open(lst, "w") as fo:
        fo.write("# list of FASTQ files\n%s\t%s\n%s\n%s\t%s\n" % (fqa[0], fqa[1], fqc[0], fqb[0], fqb[1]))
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "clil_" + tag)
        cmd = [exe, "align", "--fq_list", lst, "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "4", "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    assert sum(1 for _ in open(outs["ref"] + ".FASTQ.csv")) == 4          # header + three files
    _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")


def test_cli_two_reference_sized_batches(small_index, ref_required):
    """300,000 pairs = one full 262,144-pair batch plus a partial one at the reference's real batch size: the drand48
    position, last_ii and every accumulator cross the batch boundary exactly as in the reference (about 40 s of CPU for
    the reference run)."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    arrs = small_index.reads(300000, read_len=100, seed=95)
    fq = small_index.write_fastq("clibig", arrs)
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(small_index.dir, "clibig_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--t", str(os.cpu_count() or 4),
               "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
    recs = _compare_bams(outs["ref"] + ".bam", outs["b200"] + ".bam")
    assert len(recs) > 590000
    for tag in outs:                                   # half a gigabyte of outputs: clean up
        for ext in ("bam", "InsertSizeTable"):
            os.remove(outs[tag] + "." + ext)
    for f in fq:
        os.remove(f)


def test_cli_input_containers_give_the_same_files(small_index):
    """Row f2 at the CLI level: the same reads as a gzip stream, as BGZF (member-parallel inflate) and as plain text go
    through different producers of the feeder and must leave identical summary files and BAM records."""
    import gzip
    from test_feeder import _bgzf
    arrs = small_index.reads(20000, read_len=100, seed=87, f_on=0.95)
    fq = small_index.write_fastq("clifmt", arrs)
    inputs = {"gzip": fq, "bgzf": [], "text": []}
    for f in fq:
        text = gzip.open(f).read()
        inputs["text"].append(f[:-3]); open(f[:-3], "wb").write(text)
        inputs["bgzf"].append(f[:-6] + ".bgzf.fq.gz"); open(inputs["bgzf"][-1], "wb").write(_bgzf(text))
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, files in inputs.items():
        out = os.path.join(small_index.dir, "clifmt_" + tag)
        cmd = [CLI, "align", "--fastq_1", files[0], "--fastq_2", files[1], "--index_prefix", idx_prefix, "--out_prefix", out, "--q", "15"]
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for tag in ("bgzf", "text"):
        for ext in TEXT_FILES:
            _compare_files(outs["gzip"] + "." + ext, outs[tag] + "." + ext)
        assert len(_compare_bams(outs["gzip"] + ".bam", outs[tag] + ".bam")) > 30000
