"""A/B timing of the memory-path variants of the fast search pass (FQB_SEARCH_VAR, a bit mask: fq_kernels.cu / SearchLane
kVar) on ONE batch in one process: the align stage (prep + width + order + search) of the same 262,144 pairs, the forms
taking turns, and every form's hit lists compared entry by entry with those of form 0.  Prints one JSON line; exit code
0 = all lists equal."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastquick_b200 import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=262144)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--variants", default="0,1,4,5", help="the first one is the reference form (0 = plain loads and stores)")
    a = ap.parse_args()
    import torch
    lib = _abi.load_library()
    cfg = _abi.SynthRefCfg()
    lib.fqb_synth_ref_cfg_default(C.byref(cfg))
    s = C.c_void_p()
    assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual, g.kmer_thresh = 15, 3
    h = C.c_void_p()
    assert lib.fqb_create_from_synth(s, C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
    rc_ = _abi.SynthReadCfg()
    lib.fqb_synth_read_cfg_default(C.byref(rc_))
    rc_.read_len = a.read_len
    n, L = a.pairs, a.read_len
    arrs = [np.zeros((n, L), np.uint8) for _ in range(4)]
    assert lib.fqb_synth_reads(s, C.byref(rc_), C.c_int64(0), C.c_int64(n), *[_abi.u8p(x) for x in arrs], 0) == 0
    dev = [torch.from_numpy(x).cuda() for x in arrs]
    lib.fqb_stream.restype = C.c_void_p
    stream = torch.cuda.ExternalStream(lib.fqb_stream(h))
    ptr = [C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8)) for t in dev]
    assert lib.fqb_stage_load(h, n, L, ptr[0], ptr[1], None, ptr[2], ptr[3], None, 1) == 0, lib.fqb_last_error()

    def run(var):
        os.environ["FQB_SEARCH_VAR"] = str(var)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            rc = lib.fqb_stage_align(h)
            e1.record(stream)
        assert rc == 0, lib.fqb_last_error()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def hits():
        na = np.zeros(2 * n, np.int32)
        aln = np.zeros((2 * n, 8), _abi.ALN_DTYPE)
        assert lib.fqb_stage_fetch_aln(h, 8, aln.ctypes.data_as(C.c_void_p), _abi.i32p(na)) == 0
        keep = np.arange(8)[None, :] < np.clip(na, 0, 8)[:, None]       # entries past n_aln are not defined
        return na, aln[keep]

    variants = [int(v) for v in a.variants.split(",")]
    equal = {}
    ref = None
    for v in variants:                                                   # warm-up of every form + its hit lists
        run(v)
        got = hits()
        if ref is None:
            ref = got
        equal[v] = bool(np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]))
    ms = {v: [] for v in variants}
    for _ in range(a.rounds):
        for v in variants:
            ms[v].append(run(v))
    med = {v: float(np.median(ms[v])) for v in variants}
    ok = [v for v in variants if equal[v]]
    best = min(ok, key=lambda v: med[v])
    print(json.dumps({"tool": "stage_ab", "pairs": n, "read_len": L, "rounds": a.rounds,
                      "median_ms": {str(v): round(med[v], 3) for v in variants},
                      "all_ms": {str(v): [round(x, 3) for x in ms[v]] for v in variants},
                      "hit_lists_equal_to_form_0": {str(v): equal[v] for v in variants},
                      "best": best, "best_over_plain": med[best] / med[variants[0]],
                      "reads_with_hits": int((ref[0] > 0).sum())}), flush=True)
    lib.fqb_destroy(h)
    sys.exit(0 if all(equal.values()) else 1)


if __name__ == "__main__":
    main()
