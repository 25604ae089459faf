"""Feeder throughput (row f2), host only: records/s of fqb_feeder_fill for the same FASTQ as plain text, a gzip stream and BGZF.
usage: python tools/feeder_bench.py [n_records] [threads] [packed]     (packed: fqb_feeder_fill_packed, both forms written)"""
import ctypes as C, gzip, os, struct, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fx
from test_feeder import _bgzf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
packed = len(sys.argv) > 3 and sys.argv[3] == "packed"
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
arrs = idx.reads(n, read_len=100, seed=5)
fq = idx.write_fastq("feedbench", arrs)
text = gzip.open(fq[0]).read()
d = os.path.join(idx.dir, "feedbench")
os.makedirs(d, exist_ok=True)
paths = {"text": os.path.join(d, "r.fq"), "gzip": fq[0], "bgzf": os.path.join(d, "r.bgzf.fq.gz")}
open(paths["text"], "wb").write(text)
open(paths["bgzf"], "wb").write(_bgzf(text))
lib = fx.host_lib()
lib.fqb_feeder_fill.restype = C.c_int64
lib.fqb_feeder_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
lib.fqb_feeder_close.argtypes = [C.c_void_p]
lib.fqb_feeder_fill_packed.restype = C.c_int64
lib.fqb_feeder_fill_packed.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
cap = 262144
pk = np.zeros((cap, 32), np.uint8); qf = np.zeros((cap, 100), np.uint8)
b = np.zeros((cap, 100), np.uint8); q = np.zeros((cap, 100), np.uint8); l = np.zeros(cap, np.int32); nm = np.zeros((cap, 64), np.uint8)
for kind, p in paths.items():
    best = 0
    for rep in range(3):
        f = C.c_void_p()
        assert lib.fqb_feeder_open(p.encode(), threads, C.byref(f)) == 0
        t = time.time(); tot = 0
        while True:
            if packed:
                k = lib.fqb_feeder_fill_packed(f, cap, 100, b.ctypes.data, q.ctypes.data, l.ctypes.data, nm.ctypes.data, 64, 32, pk.ctypes.data, qf.ctypes.data)
            else:
                k = lib.fqb_feeder_fill(f, cap, 100, b.ctypes.data, q.ctypes.data, l.ctypes.data, nm.ctypes.data, 64)
            assert k >= 0, lib.fqb_last_error()
            if k == 0: break
            tot += k
        dt = time.time() - t
        lib.fqb_feeder_close(f)
        assert tot == n
        best = max(best, n / dt)
    print("%-5s %8.0f records/s  (%.0f MB/s of text, file %d MB)" % (kind, best, best * len(text) / n / 1e6, os.path.getsize(p) >> 20), flush=True)
