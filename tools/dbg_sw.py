import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fx
from fastquick_b200 import _abi
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
a = idx.reads(3000, read_len=100, seed=83, f_on=0.9)
b = idx.reads(1500, read_len=100, seed=84, sub_rate=0.03, ins_rate=0.006, del_rate=0.006, max_indel_len=3)
arrs = [np.concatenate([x, y]) for x, y in zip(a, b)]
rng = np.random.default_rng(7)
junk = rng.choice(len(arrs[0]), 300, replace=False)
for i in junk:
    e = 0 if i % 2 else 2
    arrs[e][i] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, arrs[e].shape[1])]
lib = fx.host_lib()
g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
h = C.c_void_p(); assert lib.fqb_create(idx.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0
n, L = arrs[0].shape
rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
assert lib.fqb_align_pairs(h, n, L, *[_abi.u8p(x) if x is not None else None for x in (arrs[0], arrs[1])], None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None,
                           rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0, lib.fqb_last_error()
def cig(r):
    return "".join("%d%s" % (c & 0x3fff, "MIDS"[c >> 14]) for c in r["cigar"][: r["n_cigar"]]) if r["has_cigar"] else "-"
for e in (0, 1):
    r = rows[e][3494]
    print("end", e, "pos", r["pos"], "type", r["type"], "cigar", cig(r), "mm/go/ge", r["n_mm"], r["n_gapo"], r["n_gape"])
import zlib
print("crc", zlib.crc32(rows[0].tobytes()), zlib.crc32(rows[1].tobytes()))
