"""Row f2: the parallel FASTQ feeder (fqb_feeder_*) against a plain Python parse of the same records, for the three
container formats (text, gzip stream incl. multi-member, BGZF), odd batch sizes, CRLF, a missing final newline, blank
lines, names with comments and /1 suffixes, and its error reporting."""
import ctypes as C
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi


def _records(n, seed, max_len=100, min_len=35):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        seq = bytes(np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, L)])
        qual = bytes((rng.integers(33, 74, L)).astype(np.uint8))
        tail = [b"", b"/1", b"/2", b" 1:N:0:ACGT", b"\tcomment", b"/1 extra"][i % 6]
        recs.append((b"r%d_%d" % (seed, i), tail, seq, qual))
    return recs


def _text(recs, eol=b"\n", final_newline=True, blank_tail=0):
    out = bytearray()
    for k, (name, tail, seq, qual) in enumerate(recs):
        out += b"@" + name + tail + eol + seq + eol + b"+" + eol + qual
        if k + 1 < len(recs) or final_newline: out += eol
    out += eol * blank_tail
    return bytes(out)


def _bgzf(data, block=0xff00):
    out = bytearray()
    for off in list(range(0, len(data), block)) + [None]:
        chunk = b"" if off is None else data[off:off + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        out += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, len(comp) + 25)
        out += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out)


def _read_all(path, stride, batch, n_threads=4, name_stride=64):
    lib = fx.host_lib()
    lib.fqb_feeder_fill.restype = C.c_int64
    lib.fqb_feeder_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.fqb_feeder_close.argtypes = [C.c_void_p]
    lib.fqb_feeder_format.argtypes = [C.c_void_p]
    f = C.c_void_p()
    assert lib.fqb_feeder_open(path.encode(), n_threads, C.byref(f)) == 0, lib.fqb_last_error()
    fmt = lib.fqb_feeder_format(f)
    got = []
    try:
        while True:
            b = np.zeros((batch, stride), np.uint8); q = np.zeros((batch, stride), np.uint8)
            l = np.zeros(batch, np.int32); nm = np.zeros((batch, name_stride), np.uint8)
            n = lib.fqb_feeder_fill(f, batch, stride, b.ctypes.data, q.ctypes.data, l.ctypes.data, nm.ctypes.data, name_stride)
            if n < 0: raise RuntimeError(lib.fqb_last_error().decode())
            if n == 0: break
            for i in range(n):
                got.append((bytes(nm[i]).rstrip(b"\0"), bytes(b[i]), bytes(q[i]), int(l[i])))
    finally:
        lib.fqb_feeder_close(f)
    return fmt, got


def _expect(recs, stride):
    return [(name, seq + b"N" * (stride - len(seq)), qual + b"!" * (stride - len(qual)), len(seq)) for name, _, seq, qual in recs]


@pytest.fixture(scope="module")
def recs():
    return _records(40000, 7)          # ~7.5 MB of text: crosses the 4 MiB block size and many parse runs


@pytest.mark.parametrize("fmt", ["text", "gzip", "gzip_members", "bgzf"])
@pytest.mark.parametrize("batch", [40000, 9973])
def test_feeder_formats(tmp_path, recs, fmt, batch):
    data = _text(recs)
    p = str(tmp_path / ("r.fq" if fmt == "text" else "r.fq.gz"))
    if fmt == "text": blob = data
    elif fmt == "gzip": blob = gzip.compress(data, 1)
    elif fmt == "gzip_members": blob = b"".join(gzip.compress(data[o:o + 1500000], 1) for o in range(0, len(data), 1500000))
    else: blob = _bgzf(data)
    open(p, "wb").write(blob)
    kind, got = _read_all(p, 100, batch)
    assert kind == {"text": 0, "gzip": 1, "gzip_members": 1, "bgzf": 2}[fmt]
    assert got == _expect(recs, 100)


@pytest.mark.parametrize("variant", ["crlf", "no_final_newline", "blank_tail", "one_thread"])
def test_feeder_text_quirks(tmp_path, variant):
    recs = _records(3000, 11)
    data = _text(recs, eol=b"\r\n" if variant == "crlf" else b"\n", final_newline=variant != "no_final_newline", blank_tail=3 if variant == "blank_tail" else 0)
    p = str(tmp_path / "q.fq.gz")
    open(p, "wb").write(_bgzf(data, block=777) if variant != "one_thread" else gzip.compress(data))
    _, got = _read_all(p, 128, 1024, n_threads=0)
    assert got == _expect(recs, 128)


def test_feeder_errors(tmp_path):
    recs = _records(10, 3)
    good = _text(recs)
    cases = {
        "long": (_text(_records(5, 4, max_len=100, min_len=100)), 64, "longer"),
        "qual": (good.replace(recs[4][3], recs[4][3][:-1], 1), 100, "lengths differ"),
        "header": (good.replace(b"@" + recs[6][0], b">" + recs[6][0], 1), 100, "'@'"),
        "truncated": (good[:good.rindex(recs[-1][2]) + 10], 100, "truncated"),
    }
    for tag, (data, stride, msg) in cases.items():
        p = str(tmp_path / (tag + ".fq"))
        open(p, "wb").write(data)
        with pytest.raises(RuntimeError, match=msg):
            _read_all(p, stride, 64)
    lib = fx.host_lib()
    f = C.c_void_p()
    assert lib.fqb_feeder_open(str(tmp_path / "missing.fq").encode(), 2, C.byref(f)) != 0
    open(str(tmp_path / "bad.fq.gz"), "wb").write(gzip.compress(good)[:-40] + b"x" * 40)
    with pytest.raises(RuntimeError):
        _read_all(str(tmp_path / "bad.fq.gz"), 100, 64)


def _gunzip(blob, cap, block_bytes=0):
    lib = fx.host_lib()
    lib.fqb_gunzip.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_int64()
    rc = lib.fqb_gunzip(blob, len(blob), out.ctypes.data, cap, block_bytes, C.byref(n))
    return rc, bytes(out[:n.value])


@pytest.fixture(scope="module")
def payloads():
    rng = np.random.default_rng(5)
    p = {
        "empty": b"",
        "tiny": b"hello",
        "random": rng.integers(0, 256, 300000, dtype=np.uint8).tobytes(),               # stored blocks
        "zeros": bytes(600000),                                                         # distance 1, length 258
        "period3": b"abc" * 100000, "period7": b"abcdefg" * 60000,                      # overlapping short-distance copies
        "text": b"".join(b"line %d of some text, value=%d\n" % (i, i * i) for i in range(40000)),
        "dna": np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 1000000)].tobytes(),  # short matches at long distances
    }
    p["mixed"] = p["random"][:70000] + p["text"][:200000] + p["zeros"][:100000] + p["dna"][:300000] + p["random"][:5000]
    return p


def test_inflate_matches_zlib(payloads):
    """fqb::Inflater (the serial gzip path of the feeder) against zlib's own output: every deflate block type, code
    shape and copy case, decoded through text blocks small enough that the stop/resume and window hand-over run often."""
    for name, data in payloads.items():
        blobs = [gzip.compress(data, lvl) for lvl in (0, 1, 6, 9)]
        for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            co = zlib.compressobj(6, zlib.DEFLATED, 31, 9, strategy)
            blobs.append(co.compress(data) + co.flush())
        co = zlib.compressobj(6, zlib.DEFLATED, 31)                                     # sync flushes: empty stored blocks mid-stream
        blobs.append(b"".join(co.compress(data[o:o + 50000]) + co.flush(zlib.Z_SYNC_FLUSH) for o in range(0, len(data), 50000)) + co.flush())
        for k, blob in enumerate(blobs):
            for block_bytes in (0, 640, 4096, 65536):
                rc, out = _gunzip(blob, len(data), block_bytes)
                assert rc == 0 and out == data, (name, k, block_bytes, fx.host_lib().fqb_last_error())
    whole = b"".join(payloads.values())
    blob = b"".join(gzip.compress(d, 6) for d in payloads.values()) + bytes(512)       # members back to back + zero padding
    rc, out = _gunzip(blob, len(whole), 5000)
    assert rc == 0 and out == whole


def test_inflate_rejects_corrupt_streams(payloads):
    data = payloads["text"]
    blob = gzip.compress(data, 6)
    rng = np.random.default_rng(9)
    for _ in range(200):                                                                # single bit flips: never accepted
        b = bytearray(blob)
        b[int(rng.integers(10, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = _gunzip(bytes(b), len(data) + 100000, 4096)
        assert rc != 0
    for cut in (5, 17, 18, 30, 100, len(blob) - 9, len(blob) - 8, len(blob) - 1):       # truncations
        assert _gunzip(blob[:cut], len(data), 0)[0] != 0
    assert _gunzip(blob + b"trailing garbage", len(data), 0)[0] != 0
    assert _gunzip(blob, 1000, 0)[0] != 0                                               # output buffer too small


def test_feeder_gzip_zlib_switch(tmp_path, monkeypatch):
    """FQB_GZIP_ZLIB=1 keeps zlib's inflate for gzip streams; both paths give the same records."""
    recs = _records(20000, 13)
    p = str(tmp_path / "z.fq.gz")
    open(p, "wb").write(gzip.compress(_text(recs), 6))
    _, a = _read_all(p, 100, 7777)
    monkeypatch.setenv("FQB_GZIP_ZLIB", "1")
    _, b = _read_all(p, 100, 7777)
    assert a == b == _expect(recs, 100)


def test_feeder_bgzf_member_checks(tmp_path):
    """A BGZF member is accepted only if it inflates to exactly ISIZE bytes whose CRC-32 is the stored one."""
    recs = _records(2000, 17)
    blob = bytearray(_bgzf(_text(recs), block=20000))
    size0 = (blob[16] | blob[17] << 8) + 1                                              # first member
    for tag, at, delta in (("payload", 18 + 40, 0x10), ("crc", size0 - 8, 1), ("isize", size0 - 4, 1)):
        b = bytearray(blob)
        b[at] ^= delta
        p = str(tmp_path / (tag + ".fq.gz"))
        open(p, "wb").write(bytes(b))
        with pytest.raises(RuntimeError, match="BGZF"):
            _read_all(p, 100, 512)
    # header fields of an untrusted file: a BSIZE smaller than the member's own header + trailer, and an ISIZE beyond the
    # 64 KiB a BGZF member can hold (it sizes the output buffer), must be refused before anything is read through them
    for tag, patch in (("bsize_tiny", lambda b: b.__setitem__(slice(16, 18), bytes([3, 0]))),
                       ("bsize_no_trailer", lambda b: b.__setitem__(slice(16, 18), bytes([20, 0]))),
                       ("isize_huge", lambda b: b.__setitem__(slice(size0 - 4, size0), bytes([0xff, 0xff, 0xff, 0x7f])))):
        b = bytearray(blob)
        patch(b)
        p = str(tmp_path / (tag + ".fq.gz"))
        open(p, "wb").write(bytes(b))
        with pytest.raises(RuntimeError, match="BGZF"):
            _read_all(p, 100, 512)


def test_inflate_property():
    """zlib at any level / strategy / window size -> the feeder's inflate loop at any text-block size gives the input back."""
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    fragments = st.lists(st.binary(min_size=1, max_size=60), min_size=1, max_size=10)
    payload = st.one_of(
        st.binary(max_size=5000),
        st.builds(lambda fr, picks: b"".join(fr[i % len(fr)] for i in picks), fragments, st.lists(st.integers(0, 9), max_size=6000)))

    @settings(max_examples=150, deadline=None)
    @given(payload, st.integers(0, 9), st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]),
           st.integers(9, 15), st.integers(1, 9), st.sampled_from([0, 640, 1111, 40000]))
    def check(data, level, strategy, wbits, mem, block_bytes):
        co = zlib.compressobj(level, zlib.DEFLATED, 16 + wbits, mem, strategy)
        blob = co.compress(data) + co.flush()
        rc, out = _gunzip(blob, len(data), block_bytes)
        assert rc == 0 and out == data

    check()
