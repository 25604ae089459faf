"""Experiment: several engine handles on ONE GPU, one host thread each, batches dealt round-robin with the
cross-batch state handed thread to thread (the multi-GPU protocol inside one device).  Not the contract bench."""
import argparse, ctypes as C, os, queue, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastquick_b200 import _abi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--handles", type=int, default=2)
    ap.add_argument("--pairs", type=int, default=262144)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--warm", type=int, default=4)
    a = ap.parse_args()
    import torch
    lib = _abi.load_library()
    cfg = _abi.SynthRefCfg(); lib.fqb_synth_ref_cfg_default(C.byref(cfg))
    s = C.c_void_p(); assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0
    g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
    hs = []
    for _ in range(a.handles):
        h = C.c_void_p(); assert lib.fqb_create_from_synth(s, C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error(); hs.append(h)
    rc_ = _abi.SynthReadCfg(); lib.fqb_synth_read_cfg_default(C.byref(rc_)); rc_.read_len = 100
    n, L = a.pairs, 100
    dev = []
    for b in range(a.batches):
        arrs = [np.zeros((n, L), np.uint8) for _ in range(4)]
        assert lib.fqb_synth_reads(s, C.byref(rc_), C.c_int64(b * n), C.c_int64(n), *[_abi.u8p(x) for x in arrs], 0) == 0
        dev.append([torch.from_numpy(x).cuda() for x in arrs])
    torch.cuda.synchronize()
    ptr = lambda t: C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8))
    W = a.handles
    qs = [queue.Queue() for _ in range(W)]

    def worker(t, lo, hi):
        h = hs[t]
        for b in range(lo, hi):
            if b % W != t: continue
            d = dev[b]
            assert lib.fqb_stage_load(h, n, L, ptr(d[0]), ptr(d[1]), None, ptr(d[2]), ptr(d[3]), None, 1) == 0
            assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()
            if b > lo and W > 1:
                calls, ii = qs[t].get()
                lib.fqb_set_stream_state(h, C.c_uint64(calls), C.byref(ii))
            assert lib.fqb_stage_pair(h) == 0, lib.fqb_last_error()
            if W > 1:
                calls, ii = C.c_uint64(0), _abi.ISize()
                lib.fqb_get_stream_state(h, C.byref(calls), C.byref(ii))
                qs[(t + 1) % W].put((calls.value, ii))
            assert lib.fqb_stage_sw_refine(h) == 0, lib.fqb_last_error()

    def run(lo, hi):
        for q in qs:
            while not q.empty(): q.get()
        for h in hs: lib.fqb_reset_stream(h)
        th = [threading.Thread(target=worker, args=(t, lo, hi)) for t in range(W)]
        t0 = time.time()
        for x in th: x.start()
        for x in th: x.join()
        torch.cuda.synchronize()
        return time.time() - t0

    run(0, a.warm)
    dt = run(a.warm, a.batches)
    nb = a.batches - a.warm
    print("handles %d: %.2f ms/batch  %.3e pairs/s" % (W, dt / nb * 1e3, nb * n / dt), flush=True)
    rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
    last = (a.batches - 1) % W
    assert lib.fqb_stage_fetch_rows(hs[last], rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p), None) == 0
    import zlib
    print("rows crc", zlib.crc32(rows[0].tobytes()), zlib.crc32(rows[1].tobytes()))


if __name__ == "__main__":
    main()
