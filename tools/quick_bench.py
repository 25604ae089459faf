"""Development timing of the stage kernels on one GPU (not the contract bench; see bench.py)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastquick_b200 import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=262144)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--markers", type=int, default=10197)
    ap.add_argument("--thresh", type=int, default=3)
    a = ap.parse_args()
    import torch
    lib = _abi.load_library()
    cfg = _abi.SynthRefCfg()
    lib.fqb_synth_ref_cfg_default(C.byref(cfg))
    if a.markers != 10197:
        cfg.n_long, cfg.n_short, cfg.n_x, cfg.n_y = a.markers // 10, a.markers - a.markers // 10, 0, 0
    s = C.c_void_p()
    t0 = time.time()
    assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual, g.kmer_thresh = 15, a.thresh
    h = C.c_void_p()
    rc = lib.fqb_create_from_synth(s, C.byref(g), None, 0, C.byref(h))
    assert rc == 0, lib.fqb_last_error()
    print("index build+upload %.1fs" % (time.time() - t0), flush=True)
    rc_ = _abi.SynthReadCfg()
    lib.fqb_synth_read_cfg_default(C.byref(rc_))
    rc_.read_len = a.read_len
    n, L = a.pairs, a.read_len
    arrs = [np.zeros((n, L), np.uint8) for _ in range(4)]
    t0 = time.time()
    assert lib.fqb_synth_reads(s, C.byref(rc_), C.c_int64(0), C.c_int64(n), *[_abi.u8p(x) for x in arrs], 0) == 0
    print("reads gen %.1fs" % (time.time() - t0), flush=True)
    dev = [torch.from_numpy(x).cuda() for x in arrs]
    lib.fqb_stream.restype = C.c_void_p
    stream = torch.cuda.ExternalStream(lib.fqb_stream(h))
    ptr = [C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8)) for t in dev]
    assert lib.fqb_stage_load(h, n, L, ptr[0], ptr[1], None, ptr[2], ptr[3], None, 1) == 0, lib.fqb_last_error()
    for it in range(a.iters):
        c0 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            rc = lib.fqb_stage_align(h)
            e1.record(stream)
        assert rc == 0, lib.fqb_last_error()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c1 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c1)
        pops, occ = c1[0] - c0[0], c1[1] - c0[1]
        print("iter %d: %.2f ms  %.3e pairs/s  pops/read %.1f occ/read %.1f overflow %d  occ-steps/s %.3e" % (
            it, ms, n / ms * 1e3, pops / (2 * n), occ / (2 * n), c1[3], occ / ms * 1e3), flush=True)
    na = np.zeros(2 * n, np.int32)
    aln = np.zeros((2 * n, 8), _abi.ALN_DTYPE)
    assert lib.fqb_stage_fetch_aln(h, 8, aln.ctypes.data_as(C.c_void_p), _abi.i32p(na)) == 0
    print("n_aln hist", np.bincount(np.clip(na, 0, 9)))
    lib.fqb_destroy(h)


if __name__ == "__main__":
    main()
