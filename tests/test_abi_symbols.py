"""The C-ABI library loads without a GPU and exports every symbol include/fastquick_b200.h declares."""
import ctypes as C
import os
import re

import fx
from fastquick_b200 import _abi


def _declared_symbols():
    hdr = open(os.path.join(fx.REPO, "include", "fastquick_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fqb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_python_mirror_agree():
    assert _declared_symbols() == sorted(_abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = fx.host_lib()
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # sizes the C side asserts on as well (fq_engine.cu static_asserts)
    assert C.sizeof(_abi.GapOpt) == 88
    assert C.sizeof(_abi.PeOpt) == 40
    assert _abi.READ_DTYPE.itemsize == 96
    assert _abi.ALN_DTYPE.itemsize == 16


def test_defaults_are_the_reference_defaults():
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    # gap_init_opt, libbwa/bwtaln.c:24-48
    assert (g.s_mm, g.s_gapo, g.s_gape, g.max_gapo, g.max_gape) == (3, 11, 4, 1, 6)
    assert (g.indel_end_skip, g.max_del_occ, g.max_entries, g.seed_len, g.max_seed_diff, g.max_top2) == (5, 10, 2000000, 32, 2, 30)
    assert g.fnr == 0.02 and g.mode == 3 and g.kmer_thresh == 3
    p = _abi.PeOpt()
    lib.fqb_pe_opt_default(C.byref(p))
    # bwa_init_pe_opt, libbwa/bwape.c:7-20
    assert (p.max_isize, p.max_occ, p.n_multi, p.N_multi, p.is_sw, p.type) == (500, 100000, 3, 10, 1, 1)
    assert p.ap_prior == 1e-5


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    lib = fx.host_lib()
    h = C.c_void_p()
    rc = lib.fqb_create(b"/nonexistent", None, None, 0, C.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in lib.fqb_last_error() or b"CUDA" in lib.fqb_last_error()
