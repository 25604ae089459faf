"""The warp-cooperative DP functions (fq_dp_warp.cuh: wavefront forward pass and global alignment of fq_dp_wave.cuh, row-chunk
reverse pass) run on the host by a 32-fiber SIMT emulation (tests/emul/fq_emul_warp.cpp) against the per-lane statements of
fq_device_dp.cuh, which tests/test_emul_golden.py pins to the reference's rows.  Random windows, reads with substitutions and
indels, both strands, Ns, band and window sizes around every dispatch boundary."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB = os.path.join(HERE, "emul", "libfqemul_warp.so")


def build():
    import glob
    src = os.path.join(HERE, "emul", "fq_emul_warp.cpp")
    deps = [src] + glob.glob(os.path.join(REPO, "fastquick_b200", "csrc", "*.cuh"))
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, src])
    lib = C.CDLL(LIB)
    lib.emulw_check.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def lib():
    return build()


def pack_pac(codes):
    n = len(codes)
    pad = (-n) % 4
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)


def mutate(rng, seq, sub, indel, max_indel=6):
    out = []
    i = 0
    while i < len(seq):
        u = rng.random()
        if u < sub:
            out.append((seq[i] + rng.integers(1, 4)) % 4)
            i += 1
        elif u < sub + indel:
            k = int(rng.integers(1, max_indel + 1))
            if rng.random() < 0.5:
                out.extend(rng.integers(0, 4, k))
            else:
                i += k
        else:
            out.append(seq[i])
            i += 1
    return np.array(out, np.uint8)


def make_read(rng, genome, start, length, sub, indel, strand, n_frac=0.0):
    src = genome[start:start + length + 40]
    rd = mutate(rng, src, sub, indel)[:length]
    if len(rd) < length:
        rd = np.concatenate([rd, rng.integers(0, 4, length - len(rd)).astype(np.uint8)])
    if strand:                       # fwd = the read as sequenced; alignment orientation is its reverse complement
        rd = (3 - rd[::-1]).astype(np.uint8)
    rd = rd.copy()
    if n_frac:
        rd[rng.random(length) < n_frac] = 4
    return np.ascontiguousarray(rd)


def sw_sizes(L):
    return max(2 * 704, 17 * (L + 1)), 704, 768


def refine_sizes(L):
    wcols = L + 17
    ints = max(6 * wcols, (17 if wcols <= 129 else 33) * (L + 1))
    return ints, (wcols + 3) & ~3, (wcols + L + 2 + 3) & ~3


def run(lib, mode, pac, l_pac, beg, l, rd, strand, gap_end=-1, band=50, ext=0, sizes=None):
    info = (C.c_int * 8)()
    ints, refc, ops = sizes
    rc = lib.emulw_check(mode, pac.ctypes.data, l_pac, beg, l, rd.ctypes.data, len(rd), strand, gap_end, band, ext, ints, refc, ops, info)
    return rc, list(info)


def test_local_alignment_windows(lib):
    rng = np.random.default_rng(20261017)
    genome = rng.integers(0, 4, 6000).astype(np.uint8)
    # a few low-complexity stretches so that ties between cells are common
    genome[1000:1200] = np.tile(np.array([0, 1], np.uint8), 100)
    genome[3000:3150] = 2
    pac = pack_pac(genome)
    n_fit = 0
    for it in range(500):
        L = int(rng.choice([36, 76, 100, 100, 100, 125, 150]))
        wl = int(rng.choice([24, 60, 130, 255, 256, 257, 300, 511, 512, 513, 600, 640, 700, 702]))
        beg = int(rng.integers(0, len(genome) - wl - 200))
        strand = int(rng.integers(0, 2))
        inside = rng.random() < 0.85
        start = beg + int(rng.integers(0, max(1, wl - L // 2))) if inside else int(rng.integers(0, len(genome) - 300))
        rd = make_read(rng, genome, start, L, sub=float(rng.choice([0.0, 0.01, 0.04, 0.1])), indel=float(rng.choice([0.0, 0.005, 0.02])), strand=strand,
                       n_frac=float(rng.choice([0, 0, 0.02])))
        rc, info = run(lib, 0, pac, len(genome), beg, wl, rd, strand, sizes=sw_sizes(L))
        assert rc in (0, -1), (it, rc, L, wl, beg, strand, info)
        n_fit += rc == 0
    assert n_fit > 450


def test_global_alignment_bands(lib):
    rng = np.random.default_rng(7)
    genome = rng.integers(0, 4, 4000).astype(np.uint8)
    genome[500:640] = np.tile(np.array([0, 0, 1], np.uint8), 47)[:140]
    pac = pack_pac(genome)
    n_fit = 0
    for it in range(700):
        L = int(rng.choice([1, 2, 5, 20, 36, 76, 100, 100, 126, 127, 128, 150, 200, 250]))
        dl = int(rng.choice([-12, -3, -1, 0, 0, 1, 2, 5, 9, 16, 40]))
        wl = max(1, L + dl)
        beg = int(rng.integers(0, len(genome) - wl - 300))
        strand = int(rng.integers(0, 2))
        rd = make_read(rng, genome, beg + max(0, dl // 2), L, sub=float(rng.choice([0.0, 0.02, 0.08])), indel=float(rng.choice([0.0, 0.01, 0.03])), strand=strand,
                       n_frac=float(rng.choice([0, 0, 0.03])))
        band = int(rng.choice([1, 3, 10, 50, 50, 100, 200]))
        gap_end = int(rng.choice([-1, 5]))
        sizes = refine_sizes(max(L, wl)) if rng.random() < 0.5 else sw_sizes(L)
        rc, info = run(lib, 1, pac, len(genome), beg, wl, rd, strand, gap_end=gap_end, band=band, sizes=sizes)
        assert rc in (0, -1), (it, rc, L, wl, beg, strand, band, gap_end, info)
        n_fit += rc == 0
    assert n_fit > 600


def test_refine_and_sw_core(lib):
    rng = np.random.default_rng(99)
    genome = rng.integers(0, 4, 5000).astype(np.uint8)
    pac = pack_pac(genome)
    n_cig = 0
    for it in range(300):
        L = int(rng.choice([76, 100, 100, 150]))
        strand = int(rng.integers(0, 2))
        pos = int(rng.integers(50, len(genome) - 400))
        rd = make_read(rng, genome, pos, L, sub=0.01, indel=0.01, strand=strand)
        ext = int(rng.integers(1, 9)) * (1 if strand else -1)
        rc, info = run(lib, 2, pac, len(genome), pos, 0, rd, strand, ext=ext, sizes=refine_sizes(L))
        assert rc in (0, -1), ("refine", it, rc, L, pos, strand, ext, info)
        n_cig += rc == 0 and info[0] > 0
        wl = int(rng.integers(150, 700))
        beg = max(0, pos - int(rng.integers(0, wl - L // 2)))
        rc, info = run(lib, 3, pac, len(genome), beg, wl, rd, strand, sizes=sw_sizes(L))
        assert rc in (0, -1), ("sw", it, rc, L, beg, wl, strand, info)
    assert n_cig > 200
