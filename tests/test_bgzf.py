"""Row f1, host side: the BAM writer's BGZF members (fq_bam.cpp over the library's one-shot deflate, fq_deflate.cpp)
are valid gzip members that zlib inflates back to the payload, carry BSIZE / CRC-32 / ISIZE as the BGZF format asks,
and decode through the feeder's own inflate loop as well."""
import ctypes as C
import gzip
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest

import fx


def _compress(data):
    lib = fx.host_lib()
    lib.fqb_bgzf_compress.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    cap = len(data) + 64 * (len(data) // 0xff00 + 1)
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_int64()
    assert lib.fqb_bgzf_compress(data, len(data), out.ctypes.data, cap, C.byref(n)) == 0, lib.fqb_last_error()
    return bytes(out[:n.value])


def _members(blob):
    off = 0
    while off < len(blob):
        magic, flg, xlen, si, slen, bsize = struct.unpack_from("<3sB6xH2sHH", blob, off)
        assert magic == b"\x1f\x8b\x08" and flg == 4 and xlen == 6 and si == b"BC" and slen == 2
        yield blob[off:off + bsize + 1]
        off += bsize + 1
    assert off == len(blob)


def _payloads():
    rng = np.random.default_rng(3)
    fib = [1, 1]
    while sum(fib) < 60000: fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib)])      # Huffman depth > 15 before limiting
    rng.shuffle(skew)
    dna = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300000)].tobytes()
    text = b"".join(b"read_%d\tchr%d\t%d\t60\t100M\t=\t%d\n" % (i, i % 22 + 1, i * 37, i * 37 + 300) for i in range(8000))
    yield "empty", b""
    for n in (1, 2, 3, 7, 8, 9, 15, 16, 17, 258, 259, 0xff00 - 1, 0xff00, 0xff00 + 1, 3 * 0xff00 + 5):
        yield "text_%d" % n, (text * 3)[:n]
        yield "zeros_%d" % n, bytes(n)
    yield "random", rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()
    yield "skewed", skew.tobytes()
    geo = np.concatenate([np.full(max(1, int(30000 * 0.93 ** i)), i, np.uint8) for i in range(220)])   # 220 symbols, geometric: the 15-bit limit binds
    rng.shuffle(geo)
    yield "geometric", geo.tobytes()[:65000]
    yield "dna", dna
    yield "period", b"ab" * 40000 + b"abcdefghijklmnopq" * 5000
    yield "far_repeat", dna[:40000] + dna[:40000]                                    # second half matches 40,000 back: beyond deflate's 32 KiB
    yield "two_symbols", bytes(rng.integers(0, 2, 70000, dtype=np.uint8))
    packed = rng.integers(0, 256, 50, dtype=np.uint8).tobytes()                      # BAM-like records: ids, packed bases, qualities
    recs = b"".join(struct.pack("<iiiBBHHHiiii", 300, i % 84, i * 11, 20, 60, 4680, 1, 99, 100, i % 84, i * 11 + 250, 350)
                    + b"read_%07d\0" % i + packed + bytes(rng.integers(2, 42, 100, dtype=np.uint8)) for i in range(3000))
    yield "bam_like", recs
    for k in range(40):                                                              # assorted alphabets and run structures
        a = int(rng.integers(1, 257)); n = int(rng.integers(1, 70000)); r = int(rng.integers(1, 40))
        yield "fuzz_%d" % k, np.repeat(rng.integers(0, a, n // r + 1, dtype=np.uint16).astype(np.uint8), r)[:n].tobytes()


def test_bgzf_members_round_trip():
    lib = fx.host_lib()
    lib.fqb_gunzip.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    for name, data in _payloads():
        blob = _compress(data)
        assert (gzip.decompress(blob) if blob else b"") == data, name
        sizes = []
        for m in _members(blob):
            raw = zlib.decompress(m[18:-8], -15)                                     # a complete raw deflate stream, nothing behind it
            crc, isize = struct.unpack("<II", m[-8:])
            assert crc == zlib.crc32(raw) and isize == len(raw) and len(m) <= 65536, name
            sizes.append(len(raw))
        assert sizes == [0xff00] * (len(data) // 0xff00) + ([len(data) % 0xff00] if len(data) % 0xff00 else []), name
        assert len(blob) <= len(data) + 31 * len(sizes), name                        # never worse than a stored block
        out = np.zeros(max(len(data), 1), np.uint8)
        n = C.c_int64()
        assert lib.fqb_gunzip(blob, len(blob), out.ctypes.data, len(data), 1000, C.byref(n)) == 0, (name, lib.fqb_last_error())
        assert bytes(out[:n.value]) == data, name


def test_bgzf_compresses_like_zlib_level_1():
    text = b"".join(b"read_%d\tchr%d\t%d\t60\t100M\t=\t%d\tACGTTGCA\n" % (i, i % 22 + 1, i * 37, i * 37 + 300) for i in range(20000))
    ours = len(_compress(text))
    theirs = sum(len(zlib.compress(text[o:o + 0xff00], 1)) + 20 for o in range(0, len(text), 0xff00))
    assert ours < 1.15 * theirs


def test_bgzf_zlib_level_switch():
    """FQB_BAM_LEVEL=6 puts zlib back (read once per process, hence the child process)."""
    code = ("import sys; sys.path[:0] = [%r, %r]; import gzip, test_bgzf as t; d = b'some text to compress, ' * 9000; "
            "b = t._compress(d); assert gzip.decompress(b) == d; print(len(b))") % (fx.REPO, os.path.join(fx.REPO, "tests"))
    sizes = {}
    for lvl in (None, "6", "0"):
        env = dict(os.environ)
        env.pop("FQB_BAM_LEVEL", None)
        if lvl is not None: env["FQB_BAM_LEVEL"] = lvl
        sizes[lvl] = int(subprocess.check_output([sys.executable, "-c", code], env=env).split()[-1])
    assert sizes["0"] > 9000 * 23 and sizes["6"] < sizes["0"] // 20 and sizes[None] < sizes["0"] // 20


@pytest.mark.parametrize("owned", [0, 1])
def test_bgzf_writer_file(tmp_path, owned):
    """BgzfWriter end to end (queue, writer thread, compression helpers): the file inflates to what was written, every
    member holds at most 0xff00 bytes, and the last member is the empty end-of-file block."""
    rng = np.random.default_rng(8)
    data = b"".join(b"read_%07d\t%d\t" % (i, i * 13) + bytes(rng.integers(33, 74, 100, dtype=np.uint8)) + b"\n" for i in range(150000))
    lib = fx.host_lib()
    lib.fqb_bgzf_write_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int32]
    for piece in (1 << 22, 100003, 977 if owned == 0 else 300007):
        p = str(tmp_path / ("w%d_%d.bgzf" % (owned, piece)))
        assert lib.fqb_bgzf_write_file(p.encode(), data, len(data), piece, owned) == 0, lib.fqb_last_error()
        blob = open(p, "rb").read()
        assert gzip.decompress(blob) == data
        members = list(_members(blob))
        assert members[-1] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
        assert all(struct.unpack("<I", m[-4:])[0] <= 0xff00 for m in members)


def test_bgzf_round_trip_property():
    """Any byte string survives compress -> zlib inflate and compress -> the feeder's inflate loop (hypothesis draws
    short and long inputs, low-entropy alphabets and repeated fragments)."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    lib = fx.host_lib()
    lib.fqb_gunzip.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    fragments = st.lists(st.binary(min_size=1, max_size=40), min_size=1, max_size=12)
    payload = st.one_of(
        st.binary(max_size=3000),
        st.builds(lambda fr, picks: b"".join(fr[i % len(fr)] for i in picks), fragments, st.lists(st.integers(0, 11), max_size=4000)),
        st.builds(lambda alphabet, picks: bytes(alphabet[i % len(alphabet)] for i in picks), st.binary(min_size=1, max_size=6),
                  st.lists(st.integers(0, 5), max_size=20000)))

    @settings(max_examples=150, deadline=None)
    @given(payload)
    def check(data):
        blob = _compress(data)
        assert (gzip.decompress(blob) if blob else b"") == data
        out = np.zeros(max(len(data), 1), np.uint8)
        n = C.c_int64()
        assert lib.fqb_gunzip(blob, len(blob), out.ctypes.data, len(data), 700, C.byref(n)) == 0
        assert bytes(out[:n.value]) == data

    check()
