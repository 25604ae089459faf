"""BAM writer throughput (row f1), host only: MB/s of BgzfWriter (writer thread + compression helpers) on FASTQ-like
payload, with the library's own deflate and with zlib at FQB_BAM_LEVEL.
usage: python tools/bgzf_bench.py [megabytes]"""
import ctypes as C, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 2 and sys.argv[2] == "child":
    import numpy as np
    import fx
    mb = int(sys.argv[1])
    rng = np.random.default_rng(1)
    n = mb * 1000000 // 230
    lut = np.frombuffer(b"ACGT", np.uint8)
    recs = []
    for i in range(2000):
        recs.append(b"@SIM.%d %d/1\n" % (i, i) + lut[rng.integers(0, 4, 100)].tobytes() + b"\n+\n" +
                    np.frombuffer(b"#-27<AFJ", np.uint8)[np.clip(rng.normal(6, 1.5, 100).astype(int), 0, 7)].tobytes() + b"\n")
    data = b"".join(recs[int(k)] for k in rng.integers(0, 2000, n))
    lib = fx.host_lib()
    lib.fqb_bgzf_write_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int32]
    with tempfile.TemporaryDirectory() as d:
        best = 0
        for rep in range(3):
            p = os.path.join(d, "o.bgzf")
            t = time.time()
            assert lib.fqb_bgzf_write_file(p.encode(), data, len(data), 1 << 22, 1) == 0
            best = max(best, len(data) / (time.time() - t) / 1e6)
            size = os.path.getsize(p)
    print("%-8s %7.0f MB/s  ratio %.3f" % (os.environ.get("FQB_BAM_LEVEL", "own"), best, size / len(data)))
else:
    mb = sys.argv[1] if len(sys.argv) > 1 else "200"
    for lvl in (None, "1", "6"):
        env = dict(os.environ); env.pop("FQB_BAM_LEVEL", None)
        if lvl: env["FQB_BAM_LEVEL"] = lvl
        subprocess.check_call([sys.executable, os.path.abspath(__file__), mb, "child"], env=env)
