"""Wall-clock of `FASTQuick_b200 align` (and optionally the reference CLI) on a FASTQ pair of N synthetic pairs, for the input
containers the feeder distinguishes (gzip stream, BGZF, plain text), with and without BAM output.
usage: python tools/cli_throughput.py [n_pairs] [--ref]"""
import gzip, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fx
from test_feeder import _bgzf
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 524288
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
t0 = time.time(); arrs = idx.reads(n, read_len=100, seed=99); fq = idx.write_fastq("thr", arrs)
inputs = {"gzip": fq, "bgzf": [], "text": []}
for f in fq:
    text = gzip.open(f).read()
    inputs["text"].append(f[:-3]); open(f[:-3], "wb").write(text)
    inputs["bgzf"].append(f[:-6] + ".bgzf.fq.gz"); open(inputs["bgzf"][-1], "wb").write(_bgzf(text))
print("fastq written %.1fs" % (time.time() - t0), flush=True)
CLI = os.path.join(fx.REPO, "fastquick_b200", "FASTQuick_b200")
runs = [("b200 %s%s" % (k, " +bam" if bam else ""), CLI, inputs[k], [] if bam else ["--sam_out"]) for k in ("gzip", "bgzf", "text") for bam in (False, True)]
if "--emit-modes" in sys.argv:          # host phases of the emit calls inline (default) against FQB_ASYNC_EMIT=1 (own threads), BAM runs only
    runs = [(tag + (" sync-emit" if sync else " async-emit"), exe, files, extra + (["#sync"] if sync else [])) for tag, exe, files, extra in runs if "+bam" in tag for sync in (True, False)]
if "--ref" in sys.argv: runs.append(("ref gzip +bam", fx.REF_BIN, fq, []))
for tag, exe, files, extra in runs:
    out = os.path.join(idx.dir, "thr_out")
    env = dict(os.environ)
    if "--emit-modes" in sys.argv and "#sync" not in extra: env["FQB_ASYNC_EMIT"] = "1"
    cmd = [exe, "align", "--fastq_1", files[0], "--fastq_2", files[1], "--index_prefix", idx.prefix[:-len(".FASTQuick.fa")], "--out_prefix", out, "--t", str(os.cpu_count()), "--q", "15"] + [x for x in extra if x != "#sync"]
    best = None
    for rep in range(3):                 # the boxes are shared: best of three
        t0 = time.time()
        r = subprocess.run(cmd, cwd=idx.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        dt = time.time() - t0
        sec = [l for l in r.stdout.splitlines() if "Processed Pair End mapping" in l]
        t_map = float(sec[-1].split(" in ")[1].split()[0]) if sec else float("nan")
        if best is None or t_map < best[1]: best = (dt, t_map)
        if r.returncode: break
    dt, t_map = best
    print("%-16s rc %d  wall %.2fs  mapping %.2fs = %.0f pairs/s" % (tag, r.returncode, dt, t_map, n / t_map), flush=True)
    if r.returncode: print(r.stdout[-500:])
    if os.environ.get("FQB_BAM_DEBUG"):
        for l in r.stdout.splitlines():
            if "bam_emit" in l: print("     ", l)
