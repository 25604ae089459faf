import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fx
from fastquick_b200 import _abi
idx = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
arrs = idx.reads(300000, read_len=100, seed=95)
n = 262144
sub = [np.ascontiguousarray(a[:n]) for a in arrs]
lib = fx.host_lib()
g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
h = C.c_void_p(); assert lib.fqb_create(idx.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0
L = 100
assert lib.fqb_stage_load(h, n, L, _abi.u8p(sub[0]), _abi.u8p(sub[1]), None, _abi.u8p(sub[2]), _abi.u8p(sub[3]), None, 0) == 0
assert lib.fqb_stage_align(h) == 0
def show(tag):
    rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
    ii = _abi.ISize()
    assert lib.fqb_stage_fetch_rows(h, rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), C.byref(ii)) == 0
    for e in (0, 1):
        r = rows[e][39552]
        cg = "".join("%d%s" % (c & 0x3fff, "MIDS"[c >> 14]) for c in r["cigar"][: r["n_cigar"]]) if r["has_cigar"] else "-"
        print(tag, "end", e, "pos", r["pos"], "strand", r["strand"], "type", r["type"], "flag", r["extra_flag"], "mapQ", r["mapQ"], "seQ", r["seQ"], "len", r["len"], "mm/go/ge", r["n_mm"], r["n_gapo"], r["n_gape"], "cigar", cg)
    print(tag, "isize", ii.avg, ii.std, ii.low, ii.high, ii.high_bayesian)
assert lib.fqb_stage_pair(h) == 0
show("pair")
assert lib.fqb_stage_sw_refine(h) == 0
show("final")
