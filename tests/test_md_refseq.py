"""The reference's own known-answer tests for this path (test/alltests.cpp:118-151, RecoverRefseqByMDandCigar) run
against the oracle's restatement, the oracle's MD tag against the reference's bwa_cal_md1 output, and the property that
ties them to what the CUDA path does instead (a direct look-up of the packed reference): for every aligned read,
RecoverRefseq(read, MD, CIGAR) equals the pac bases under the read's M and D operations."""
import ctypes as C

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi

# (read, MD, CIGAR, expected) -- vectors of the reference's test3.c1
KATS = [
    ("AAAAAATAAAAAA", "T12", "13M", "TAAAAATAAAAAA"),
    ("AAAAAATAAAAAA", "4T8", "13M", "AAAATATAAAAAA"),
    ("AAAAAATAAAAAA", "12T", "13M", "AAAAAATAAAAAT"),
    ("AAAAAATAAAAAA", "11T0T", "13M", "AAAAAATAAAATT"),
    ("AAAAAATAAAAAA", "13", "4M3I9M", "AAAAAATAAAAAA"),
    ("AAAAAATAAAAAA", "9", "4S9M", "AATAAAAAA"),
    ("AAAAAATAAAAAA", "9", "9M4S", "AAAAAATAA"),
    ("AAAAAATAAAAAA", "9^TTTT4", "9M4D4M", "AAAAAATAATTTTAAAA"),
    ("AAAAAATAAAAAA", "6G^GGG5T", "7M3D6M", "AAAAAAGGGGAAAAAT"),
    ("AAAAAATAAAAAA", "9^TTTT2^TT2", "9M4D2M2D2M", "AAAAAATAATTTTAATTAA"),
]
OPS = {"M": 0, "I": 1, "D": 2, "S": 3}


def _cigar(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append(OPS[ch] << 14 | int(num)); num = ""
    return np.array(out, np.uint16)


def _recover(lib, read, md, cig):
    buf = C.create_string_buffer(4096)
    lib.orc_recover_refseq(read.encode(), md.encode(), cig.ctypes.data_as(C.c_void_p), len(cig), buf, 4096)
    return buf.value.decode()


@pytest.mark.parametrize("read,md,cigar,expected", KATS)
def test_reference_kats_recover_refseq(read, md, cigar, expected):
    assert _recover(fx.build_oracle(), read, md, _cigar(cigar)) == expected


def _oriented(codes, strand):
    if not strand:
        return codes
    c = codes[::-1].copy()
    m = c < 4
    c[m] = 3 - c[m]
    return c


def test_md_matches_reference_and_recovers_pac(small_index, ref_required):
    arrs = small_index.reads(1500, read_len=100, seed=91, sub_rate=0.02, ins_rate=0.004, del_rate=0.004, max_indel_len=3)
    fq = small_index.write_fastq("md", arrs)
    ref = fx.RefRun(small_index.prefix, fq[0], fq[1], trim_qual=15)
    assert ref.next_batch() == 1500
    lib = fx.build_oracle()
    pac = np.fromfile(small_index.prefix + ".pac", dtype=np.uint8)
    import oracle_py
    orc = oracle_py.Oracle(small_index.prefix)
    l_pac = int(orc.b0.seq_len)
    pacp = np.concatenate([pac, np.zeros(16, np.uint8)])
    n_checked = n_gapped = 0
    for e in (0, 1):
        rows = ref.rows(3, e)
        mds = ref.md(e)
        codes = fx.NT4[arrs[0 if e == 0 else 2]]
        for i in range(len(rows)):
            s = rows[i]
            if s["type"] == 0 or s["filtered"]:          # BWA_TYPE_NO_MATCH
                continue
            seq = np.ascontiguousarray(_oriented(codes[i][: int(s["full_len"])], int(s["strand"])))
            cig = np.ascontiguousarray(s["cigar"][: int(s["n_cigar"])])
            buf = C.create_string_buffer(1024)
            nm = lib.orc_cal_md(int(s["n_cigar"]), cig.ctypes.data_as(C.c_void_p), int(s["has_cigar"]), int(s["len"]), C.c_uint32(int(s["pos"])),
                                _abi.u8p(seq), C.c_int64(l_pac), _abi.u8p(pacp), buf, 1024)
            assert buf.value.decode() == mds[i], (e, i, buf.value, mds[i])
            assert nm == int(s["nm"])
            # inverse: the reference bases StatCollector would rebuild == the pac bases the CUDA kernels read directly
            read_txt = "".join("ACGTN"[c] for c in seq)
            got = _recover(lib, read_txt, mds[i], cig if s["has_cigar"] else np.zeros(0, np.uint16))
            want, x = [], int(s["pos"])
            ops = [(int(c) >> 14, int(c) & 0x3fff) for c in cig] if s["has_cigar"] else [(0, int(s["len"]))]
            for op, l in ops:
                if op in (0, 2):
                    for z in range(l):
                        k = x + z
                        if k < l_pac:
                            want.append("ACGT"[(int(pacp[k >> 2]) >> ((~k & 3) << 1)) & 3])
                    x += l
            # an N in the read is reported by MD with the reference base, so the rebuilt sequence is exactly pac
            assert got == "".join(want), (e, i, mds[i], got, "".join(want))
            n_checked += 1
            n_gapped += any(op == 2 for op, _ in ops)
    assert n_checked > 2000 and n_gapped > 20
