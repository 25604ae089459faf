"""Minimal BAM reader for the tests (BGZF = concatenated gzip members; BAM v1 records)."""
import gzip
import struct


def read_bam(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode()
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o); o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, o); o += 4
        name = data[o:o + l_name - 1].decode(); o += l_name
        l_ref, = struct.unpack_from("<i", data, o); o += 4
        refs.append((name, l_ref))
    recs = []
    while o < len(data):
        bs, = struct.unpack_from("<i", data, o); o += 4
        r = data[o:o + bs]; o += bs
        ref_id, pos, l_rn, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", r, 0)
        p = 32
        name = r[p:p + l_rn - 1].decode(); p += l_rn
        cig = []
        for _ in range(n_cig):
            v, = struct.unpack_from("<I", r, p); p += 4
            cig.append("%d%s" % (v >> 4, "MIDNSHP=X"[v & 15]))
        nb = (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(r[p + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)); p += nb
        qual = bytes(r[p:p + l_seq]); p += l_seq
        tags = {}
        while p < len(r):
            tag = r[p:p + 2].decode(); t = chr(r[p + 2]); p += 3
            if t in "cCsSiI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[t]
                v, = struct.unpack_from(fmt, r, p); p += struct.calcsize(fmt)
                tags[tag] = ("i", v)
            elif t == "A":
                tags[tag] = ("A", chr(r[p])); p += 1
            elif t == "Z":
                e = r.index(b"\0", p)
                tags[tag] = ("Z", r[p:e].decode()); p = e + 1
            elif t == "f":
                v, = struct.unpack_from("<f", r, p); p += 4
                tags[tag] = ("f", v)
            else:
                raise ValueError("tag type " + t)
        recs.append(dict(name=name, flag=flag, ref=refs[ref_id][0] if ref_id >= 0 else "*", pos=pos + 1, mapq=mapq, bin=bin_, cigar="".join(cig) or "*",
                         mref=refs[nref][0] if nref >= 0 else "*", mpos=npos + 1, tlen=tlen, seq=seq, qual=qual, tags=tags))
    return text, refs, recs
