import ctypes as C, os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from fastquick_b200 import _abi
lib = _abi.load_library()
cfg = _abi.SynthRefCfg(); lib.fqb_synth_ref_cfg_default(C.byref(cfg))
s = C.c_void_p(); assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0
g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
h = C.c_void_p(); assert lib.fqb_create_from_synth(s, C.byref(g), None, 0, C.byref(h)) == 0
work = tempfile.mkdtemp(); prefix = os.path.join(work, "bench.FASTQuick.fa")
assert lib.fqb_synth_write_inputs(s, work.encode()) == 0
assert lib.fqb_synth_write_index(s, os.path.join(work, "genome.fa").encode(), os.path.join(work, "dbsnp.vcf").encode(), prefix.encode(), 0) == 0
assert lib.fqb_stats_open(h, prefix.encode()) == 0, lib.fqb_last_error()
assert lib.fqb_stats_begin_file(h, os.path.join(work, "out").encode(), b"a", b"b") == 0
n, L = 262144, 100
rc_ = _abi.SynthReadCfg(); lib.fqb_synth_read_cfg_default(C.byref(rc_)); rc_.read_len = L
arrs = [np.zeros((n, L), np.uint8) for _ in range(4)]
assert lib.fqb_synth_reads(s, C.byref(rc_), C.c_int64(0), C.c_int64(n), *[_abi.u8p(x) for x in arrs], 0) == 0
for it in range(3):
    t0 = time.time()
    assert lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), None, None, None, None) == 0
    t1 = time.time()
    assert lib.fqb_stage_stats(h) == 0
    import torch; torch.cuda.synchronize()
    t2 = time.time()
    assert lib.fqb_stats_emit(h, None, 0) == 0
    t3 = time.time()
    print("align %.1f ms  stats %.1f ms  emit %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
t0 = time.time(); assert lib.fqb_stats_finish(h, os.path.join(work, "out").encode()) == 0, lib.fqb_last_error(); print("finish %.1f ms" % ((time.time() - t0) * 1e3))
