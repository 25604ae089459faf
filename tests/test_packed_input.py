"""The packed input form (north_star: 2-bit bases + not-ACGT flag in the quality byte, include/fastquick_b200.h "packed input
form"): fqb_pack_reads and the feeder's fill_packed against a numpy statement of the layout (CPU), and -- on the GPU -- the packed
calls against their ASCII counterparts on the same reads: rows must be bit-identical, including reads with N, lower-case
letters, '-' (nt4 code 5), quality-trimmed tails and lengths that differ inside a batch."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi


def vp(a):
    return C.c_void_p(a.ctypes.data)


def nt4_table():
    t = np.full(256, 4, np.uint8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        t[ch] = v
    t[ord("-")] = 5
    return t


def unpack(pk, qf, L):
    w = np.ascontiguousarray(pk).view(np.uint32)
    j = np.arange(L)
    field = (w[:, j >> 4] >> (2 * (j & 15)).astype(np.uint32)) & 3
    return np.where(qf & 0x80, 4 + field, field).astype(np.uint8), (qf & 0x7F).astype(np.uint8)


def pack(lib, b, q):
    n, L = b.shape
    lib.fqb_packed_stride.restype = C.c_int32
    ps = lib.fqb_packed_stride(L)
    pk, qf = np.zeros((n, ps), np.uint8), np.zeros((n, L), np.uint8)
    assert lib.fqb_pack_reads(C.c_int64(n), L, vp(b), vp(q), ps, vp(pk), vp(qf)) == 0, lib.fqb_last_error()
    return ps, pk, qf


def test_pack_reads_layout():
    lib = fx.host_lib()
    lib.fqb_packed_stride.restype = C.c_int32
    assert [lib.fqb_packed_stride(x) for x in (1, 64, 65, 100, 128, 129, 150, 256)] == [16, 16, 32, 32, 32, 48, 48, 64]
    rng = np.random.default_rng(5)
    for L in (36, 64, 100, 150, 256):
        b = rng.choice(np.frombuffer(b"ACGTACGTACGTacgtN-RYn.", dtype=np.uint8), size=(257, L))
        q = rng.integers(33, 127, size=(257, L)).astype(np.uint8)
        ps, pk, qf = pack(lib, b, q)
        codes, quals = unpack(pk, qf, L)
        assert (codes == nt4_table()[b]).all() and (quals == q).all()
        assert not pk.view(np.uint32)[:, (L + 15) // 16:].any()             # padding words are zero
    q[3, 7] = 200
    ps = lib.fqb_packed_stride(256)
    assert lib.fqb_pack_reads(C.c_int64(257), 256, vp(b), vp(q), ps, vp(np.zeros((257, ps), np.uint8)), vp(np.zeros((257, 256), np.uint8))) != 0
    assert lib.fqb_pack_reads(C.c_int64(1), 100, vp(b), vp(q), 16, vp(np.zeros(64, np.uint8)), vp(np.zeros(256, np.uint8))) != 0      # wrong stride


def test_feeder_fill_packed(tmp_path):
    lib = fx.host_lib()
    lib.fqb_packed_stride.restype = C.c_int32
    lib.fqb_feeder_fill_packed.restype = C.c_int64
    rng = np.random.default_rng(11)
    n, stride, ns = 5000, 120, 32
    path = str(tmp_path / "r.fq.gz")
    with gzip.open(path, "wt") as f:
        for i in range(n):
            L = int(rng.integers(30, stride + 1))
            seq = "".join(rng.choice(list("ACGTACGTNacgt-"), L))
            qual = "".join(chr(c) for c in rng.integers(33, 127, L))
            f.write("@r%d/1\n%s\n+\n%s\n" % (i, seq, qual))
    fd = C.c_void_p()
    assert lib.fqb_feeder_open(path.encode(), 4, C.byref(fd)) == 0, lib.fqb_last_error()
    ps = lib.fqb_packed_stride(stride)
    got = 0
    while True:
        cap = 1777
        b, q = np.zeros((cap, stride), np.uint8), np.zeros((cap, stride), np.uint8)
        lens, names = np.zeros(cap, np.int32), np.zeros((cap, ns), np.uint8)
        pk, qf = np.full((cap, ps), 0xEE, np.uint8), np.zeros((cap, stride), np.uint8)
        k = lib.fqb_feeder_fill_packed(fd, cap, stride, vp(b), vp(q), vp(lens), vp(names), ns, ps, vp(pk), vp(qf))
        assert k >= 0, lib.fqb_last_error()
        if k == 0:
            break
        ps2, pk2, qf2 = pack(lib, b[:k], q[:k])
        assert ps2 == ps and (pk[:k] == pk2).all() and (qf[:k] == qf2).all()
        got += k
    assert got == n
    lib.fqb_feeder_close(fd)


@pytest.mark.gpu
def test_packed_calls_equal_ascii_calls(small_index):
    lib = fx.host_lib()
    n, L = 6000, 100
    arrs = small_index.reads(n, read_len=L, seed=4242, sub_rate=0.02, ins_rate=0.004, del_rate=0.004)
    rng = np.random.default_rng(3)
    for e in (0, 2):                                           # sprinkle N, lower case and '-' over the bases
        b = arrs[e]
        m = rng.random(b.shape)
        b[m < 0.004] = ord("N")
        b[(m > 0.004) & (m < 0.006)] = ord("-")
        low = (m > 0.5) & (m < 0.6) & (b != ord("N")) & (b != ord("-"))
        b[low] = b[low] | 0x20
    lens = [np.full(n, L, np.int32), np.full(n, L, np.int32)]
    lens[0][::7] = 97
    lens[1][::5] = 99
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    ps, pk1, qf1 = pack(lib, arrs[0], arrs[1])
    _, pk2, qf2 = pack(lib, arrs[2], arrs[3])
    i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    out = {}
    for mode in ("ascii", "packed", "packed_piped"):
        h = C.c_void_p()
        assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
        rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
        r1, r2 = rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p)
        if mode == "ascii":
            rc = lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), i32(lens[0]), _abi.u8p(arrs[2]), _abi.u8p(arrs[3]), i32(lens[1]), r1, r2, None)
        elif mode == "packed":
            rc = lib.fqb_align_pairs_packed(h, n, L, ps, _abi.u8p(pk1), _abi.u8p(qf1), i32(lens[0]), _abi.u8p(pk2), _abi.u8p(qf2), i32(lens[1]), r1, r2, None)
        else:
            rc = lib.fqb_submit_pairs_packed(h, n, L, ps, _abi.u8p(pk1), _abi.u8p(qf1), i32(lens[0]), _abi.u8p(pk2), _abi.u8p(qf2), i32(lens[1]), 0)
            assert rc == 0, lib.fqb_last_error()
            rc = lib.fqb_collect_pairs(h, r1, r2)
            assert rc == 0, lib.fqb_last_error()
            rc = lib.fqb_rows_wait(h)
        assert rc == 0, lib.fqb_last_error()
        lib.fqb_destroy(h)
        out[mode] = rows
    for mode in ("packed", "packed_piped"):
        for e in (0, 1):
            assert out[mode][e].tobytes() == out["ascii"][e].tobytes(), (mode, e)
    assert int((out["ascii"][0]["type"] != 0).sum()) > 4000          # the batch did align


@pytest.mark.gpu
def test_packed_single_end_and_150_bases(small_index):
    """Single-end input (no second reads) and 150-base reads (48-byte packed rows, three 128-bit loads) in the packed form."""
    lib = fx.host_lib()
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15
    for L, single in ((100, True), (150, False), (150, True)):
        n = 3000
        arrs = small_index.reads(n, read_len=L, seed=900 + L, sub_rate=0.015, ins_rate=0.003, del_rate=0.003)
        arrs[0][::11, 5] = ord("N")
        ps, pk1, qf1 = pack(lib, arrs[0], arrs[1])
        _, pk2, qf2 = pack(lib, arrs[2], arrs[3])
        out = []
        for packed in (False, True):
            h = C.c_void_p()
            assert lib.fqb_create(small_index.prefix.encode(), C.byref(g), None, 0, C.byref(h)) == 0, lib.fqb_last_error()
            rows = [np.zeros(n, _abi.READ_DTYPE) for _ in range(2)]
            r1, r2 = rows[0].ctypes.data_as(C.c_void_p), rows[1].ctypes.data_as(C.c_void_p)
            if packed:
                rc = lib.fqb_align_pairs_packed(h, n, L, ps, _abi.u8p(pk1), _abi.u8p(qf1), None, None if single else _abi.u8p(pk2), None if single else _abi.u8p(qf2), None, r1, r2, None)
            else:
                rc = lib.fqb_align_pairs(h, n, L, _abi.u8p(arrs[0]), _abi.u8p(arrs[1]), None, None if single else _abi.u8p(arrs[2]), None if single else _abi.u8p(arrs[3]), None, r1, r2, None)
            assert rc == 0, lib.fqb_last_error()
            lib.fqb_destroy(h)
            out.append(rows)
        assert out[0][0].tobytes() == out[1][0].tobytes(), (L, single)
        assert out[0][1].tobytes() == out[1][1].tobytes(), (L, single)
        assert int((out[0][0]["type"] != 0).sum()) > n // 2
