"""One-off large parity runs of the two command lines (not part of the test suite: minutes of reference CPU time)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fx, bamio
from test_gpu_stats import TEXT_FILES, _compare_files
CLI = os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200")
if os.environ.get("FQB_BIG_INDEX") == "10k":      # the BASELINE marker set: 1000 long + 9000 short + 100 X + 97 Y flanks
    idx = fx.SynthIndex("bench10k", n_long=1000, n_short=9000, n_x=100, n_y=97, with_rollhash=True)
else:
    idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
all_cases = {
    "seed301": (1500000, dict(read_len=100, seed=301, f_on=0.97, sub_rate=0.01, ins_rate=0.0015, del_rate=0.0015, max_indel_len=3)),
    "se100": (500000, dict(read_len=100, seed=302, sub_rate=0.012, ins_rate=0.002, del_rate=0.002, max_indel_len=3)),
    "err100": (300000, dict(read_len=100, seed=203, sub_rate=0.04, ins_rate=0.01, del_rate=0.01, max_indel_len=4)),
    "off100": (400000, dict(read_len=100, seed=204, f_on=0.3, sub_rate=0.015)),
    "big100": (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, dict(read_len=100, seed=201, f_on=0.95, sub_rate=0.012, ins_rate=0.002, del_rate=0.002, max_indel_len=3)),
    "big150": (int(sys.argv[2]) if len(sys.argv) > 2 else 300000, dict(read_len=150, seed=202, sub_rate=0.02, ins_rate=0.004, del_rate=0.004, max_indel_len=4)),
}
sel = os.environ.get("FQB_BIG_CASES", "big100,big150").split(",")
cases = {k: all_cases[k] for k in sel}
for name, (n, kw) in cases.items():
    t0 = time.time()
    arrs = idx.reads(n, **kw)
    rng = np.random.default_rng(11)
    for i in rng.choice(n, n // 50, replace=False):          # 2 % half-mapped pairs
        e = 0 if i % 2 else 2
        arrs[e][i] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, arrs[e].shape[1])]
    fq = idx.write_fastq(name, arrs)
    outs = {}
    for tag, exe in (("ref", fx.REF_BIN), ("b200", CLI)):
        out = os.path.join(idx.dir, name + "_" + tag)
        cmd = [exe, "align", "--fastq_1", fq[0]] + ([] if name.startswith("se") else ["--fastq_2", fq[1]]) + ["--index_prefix", idx.prefix[:-len(".FASTQuick.fa")], "--out_prefix", out,
               "--t", str(os.cpu_count() or 4), "--q", "15"]
        t1 = time.time()
        r = subprocess.run(cmd, cwd=idx.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        print(name, tag, "%.1fs" % (time.time() - t1), flush=True)
        outs[tag] = out
    bad = []
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        try:
            _compare_files(outs["ref"] + "." + ext, outs["b200"] + "." + ext)
        except AssertionError as ex:
            bad.append((ext, str(ex)[:600]))
    ta, ra, a = bamio.read_bam(outs["ref"] + ".bam")
    tb, rb, b = bamio.read_bam(outs["b200"] + ".bam")
    nbad = 0
    if len(a) != len(b): bad.append(("bam", "record counts %d %d" % (len(a), len(b))))
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            nbad += 1
            if nbad <= 3:
                strip = lambda r: {k: v for k, v in r.items() if k not in ("qual", "seq")}
                bad.append(("bam rec %d" % i, str((strip(x), strip(y)))[:900]))
    print(name, "pairs", n, "bam records", len(a), "differing records", nbad, "file problems", len(bad), "total %.0fs" % (time.time() - t0), flush=True)
    for b_ in bad: print("   ", b_)
    for tag in outs:
        for ext in ("bam", "InsertSizeTable"):
            os.remove(outs[tag] + "." + ext)
    for f in fq: os.remove(f)
