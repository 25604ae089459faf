"""Wall-clock of `FASTQuick_b200 align` (and optionally the reference CLI) on a FASTQ pair of N synthetic pairs."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
t0 = time.time(); arrs = idx.reads(n, read_len=100, seed=99); fq = idx.write_fastq("thr", arrs); print("fastq written %.1fs" % (time.time() - t0), flush=True)
for tag, exe, extra in (("b200", os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200"), []), ("b200-nobam", os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200"), ["--sam_out"])) + ((("ref", fx.REF_BIN, []),) if "--ref" in sys.argv else ()):
    out = os.path.join(idx.dir, "thr_" + tag)
    cmd = [exe, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx.prefix[:-len(".FASTQuick.fa")], "--out_prefix", out, "--t", str(os.cpu_count()), "--q", "15"] + extra
    t0 = time.time()
    r = subprocess.run(cmd, cwd=idx.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.time() - t0
    for l in r.stdout.splitlines():
        if "bam_emit" in l: print("   ", l)
    line = [l for l in r.stdout.splitlines() if "Processed Pair End mapping" in l]
    print(tag, "rc", r.returncode, "wall %.2fs" % dt, line[-1] if line else r.stdout[-300:], "bam %.1f MB" % (os.path.getsize(out + ".bam") / 1e6 if os.path.exists(out + ".bam") else 0), flush=True)
