"""Where the fixed cost of one `FASTQuick_b200 align` run goes: the CLI's own NOTICE lines for a small input."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fx
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
arrs = idx.reads(100000, read_len=100, seed=98); fq = idx.write_fastq("startup", arrs)
CLI = os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200")
for rep in range(2):
    t0 = time.time()
    r = subprocess.run([CLI, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", idx.prefix[:-len(".FASTQuick.fa")], "--out_prefix", os.path.join(idx.dir, "startup_out"), "--q", "15"],
                       cwd=idx.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("run %d: wall %.2fs rc %d" % (rep, time.time() - t0, r.returncode))
    for l in r.stdout.splitlines():
        if "sec" in l: print("   ", l)
