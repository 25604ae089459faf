"""Per-mode step counts of the search state machine on the bench workload, from the host instantiation of the lane
(tests/emul) -- development aid for search_kernel, no GPU needed.   usage: python tools/search_stats.py [n_pairs] [read_len]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fx
from fastquick_b200 import _abi
import oracle_py

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
t0 = time.time()
idx = fx.SynthIndex("bench10k", n_long=1000, n_short=9000, n_x=100, n_y=97, with_rollhash=False)
print("index %.1fs" % (time.time() - t0))
arrs = idx.reads(n, read_len=L)
orc = oracle_py.Oracle(idx.prefix)
lib = fx.build_emul()
lib.emul_open.restype = C.c_void_p
err = C.create_string_buffer(256)
h = C.c_void_p(lib.emul_open(idx.prefix.encode(), err, 256))
assert h, err.value
g = _abi.GapOpt(); fx.host_lib().fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
codes = np.zeros((2 * n, L), np.uint8); codes[0::2] = fx.NT4[arrs[0]]; codes[1::2] = fx.NT4[arrs[2]]
quals = np.zeros((2 * n, L), np.uint8); quals[0::2] = arrs[1]; quals[1::2] = arrs[3]
lens = np.array([orc.lib.orc_trim_len(15, _abi.u8p(np.ascontiguousarray(quals[r])), L) for r in range(2 * n)], np.int32)
out = np.zeros((2 * n, 8), _abi.ALN_DTYPE); na = np.zeros(2 * n, np.int32); st = np.zeros(2 * n, np.int32); po = np.zeros((2 * n, 2), np.uint32)
t0 = time.time()
rc = lib.emul_align(h, C.byref(g), 2 * n, L, _abi.u8p(codes), _abi.i32p(lens), 4096, 8, out.ctypes.data_as(C.c_void_p), _abi.i32p(na), _abi.i32p(st),
                    po.ctypes.data_as(C.POINTER(C.c_uint32)))
print("emul %.1fs rc %d" % (time.time() - t0, rc))
s = (C.c_ulonglong * 16)(); lib.emul_stats(s)
names = ["iter", "mempop", "skip", "exact", "expand", "push", "hit", "top", "reads", "adiff", "gapok", "am"]
R = 2 * n
for i, nm in enumerate(names):
    print("%-8s %12d  %.1f/read" % (nm, s[i], s[i] / R))
print("pops/read %.1f occ/read %.1f; status!=1: %d; n_aln hist %s" % (po[:, 0].mean(), po[:, 1].mean(), (st != 1).sum(), np.bincount(np.clip(na, 0, 9))))
