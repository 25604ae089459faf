"""Python driver for the plain-C oracle (oracle/fq_oracle.c).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

import fx
from fastquick_b200 import _abi


class OrcBwt(C.Structure):
    _fields_ = [("primary", C.c_uint32), ("L2", C.c_uint32 * 5), ("seq_len", C.c_uint32),
                ("bwt", C.POINTER(C.c_uint32)), ("sa_intv", C.c_uint32), ("n_sa", C.c_uint32),
                ("sa", C.POINTER(C.c_uint32))]


class OrcGapOpt(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("s_mm", "s_gapo", "s_gape", "mode", "indel_end_skip", "max_del_occ", "max_entries",
                                       "max_diff", "max_gapo", "max_gape", "max_seed_diff", "seed_len", "max_top2")]


def _load_bwt(bwt_path, sa_path):
    raw = np.fromfile(bwt_path, dtype="<u4")
    sa_raw = np.fromfile(sa_path, dtype="<u4")
    b = OrcBwt()
    b.primary = int(raw[0])
    for i in range(4):
        b.L2[i + 1] = int(raw[1 + i])
    b.seq_len = b.L2[4]
    words = np.ascontiguousarray(raw[5:])
    assert sa_raw[0] == raw[0] and sa_raw[6] == b.seq_len
    b.sa_intv = int(sa_raw[5])
    sa = np.concatenate([np.array([0xFFFFFFFF], "<u4"), sa_raw[7:]])
    b.n_sa = len(sa)
    b.bwt = words.ctypes.data_as(C.POINTER(C.c_uint32))
    b.sa = sa.ctypes.data_as(C.POINTER(C.c_uint32))
    return b, (words, sa)


class OrcRng(C.Structure):
    _fields_ = [("x", C.c_uint64), ("n_calls", C.c_uint64)]


class OrcPeOpt(C.Structure):
    _fields_ = [("max_isize", C.c_int32), ("force_isize", C.c_int32), ("max_occ", C.c_uint32), ("n_multi", C.c_int32),
                ("N_multi", C.c_int32), ("type", C.c_int32), ("is_sw", C.c_int32), ("ap_prior", C.c_double)]


class Oracle:
    def __init__(self, prefix):
        self.lib = fx.build_oracle()
        L = self.lib
        L.orc_stack_new.restype = C.c_void_p
        L.orc_occ.restype = C.c_uint32
        L.orc_sa.restype = C.c_uint32
        L.orc_drand48.restype = C.c_double
        self.prefix = prefix
        self.b0, self._k0 = _load_bwt(prefix + ".bwt", prefix + ".sa")
        self.b1, self._k1 = _load_bwt(prefix + ".rbwt", prefix + ".rsa")
        self.bwts = (C.POINTER(OrcBwt) * 2)(C.pointer(self.b0), C.pointer(self.b1))
        self._roll = None
        self.rng = OrcRng()
        self.lib.orc_srand48(C.byref(self.rng), C.c_long(11))       # srand48(bns->seed), src/BwtMapper.cpp:1817
        self.last_ii = _abi.ISize()
        self.last_ii.avg = -1.0

    def pe_batch(self, lens, full_lens, filt, out, na, cap=8, want_multi=False):
        """bwa_cal_pac_pos_pe over one batch: returns (rows[2n] as READ_DTYPE, isize, multi_pos)."""
        n2 = len(lens)
        rows = np.zeros(n2, _abi.READ_DTYPE)
        rows["len"] = lens; rows["full_len"] = full_lens; rows["clip_len"] = lens; rows["filtered"] = filt
        rows["n_aln"] = np.where(filt != 0, 0, np.minimum(na, 65535))
        _, g = self.gap_opt()
        po = OrcPeOpt()
        pe = _abi.PeOpt(); fx.host_lib().fqb_pe_opt_default(C.byref(pe))
        for n, _ in OrcPeOpt._fields_:
            setattr(po, n, getattr(pe, n))
        ii = _abi.ISize()
        multi = np.zeros((n2, 11), np.uint32) if want_multi else None
        na32 = np.ascontiguousarray(np.where(filt != 0, 0, na).astype(np.int32))
        rc = self.lib.orc_cal_pac_pos_pe(self.bwts, n2 // 2, rows.ctypes.data_as(C.c_void_p), _abi.i32p(na32),
                                         out.ctypes.data_as(C.c_void_p), cap, C.c_double(g.fnr), g.max_diff, g.s_mm,
                                         C.byref(po), C.byref(self.rng), C.byref(self.last_ii), C.byref(ii),
                                         multi.ctypes.data_as(C.c_void_p) if want_multi else None)
        assert rc == 0
        self.last_ii = ii
        return rows, ii, multi

    def pac(self):
        if not hasattr(self, "_pac"):
            self._pac = np.fromfile(self.prefix + ".pac", dtype=np.uint8)
            self._pac = np.concatenate([self._pac, np.zeros(16, np.uint8)])
        return self._pac

    def sw_and_refine(self, rows, ii, arrs):
        """bwa_paired_sw then bwa_refine_gapped on the rows of pe_batch(); returns (rows after SW, rows after refine)."""
        n, rl = arrs[0].shape
        codes = np.zeros((2 * n, rl), np.uint8)
        codes[0::2] = fx.NT4[arrs[0]]; codes[1::2] = fx.NT4[arrs[2]]
        codes = np.ascontiguousarray(codes)
        po = OrcPeOpt()
        pe = _abi.PeOpt(); fx.host_lib().fqb_pe_opt_default(C.byref(pe))
        for nme, _ in OrcPeOpt._fields_:
            setattr(po, nme, getattr(pe, nme))
        l_pac = int(self.b0.seq_len)
        pac = self.pac()
        rows = rows.copy()
        self.lib.orc_paired_sw(C.c_int64(l_pac), _abi.u8p(pac), n, rows.ctypes.data_as(C.c_void_p), _abi.u8p(codes), rl, C.byref(po), C.byref(ii))
        after_sw = rows.copy()
        self.lib.orc_refine_gapped_batch(C.c_int64(l_pac), _abi.u8p(pac), 2 * n, rows.ctypes.data_as(C.c_void_p), _abi.u8p(codes), rl)
        return after_sw, rows

    def roll_tables(self):
        if self._roll is None:
            self._roll = np.memmap(self.prefix + ".rollhash", dtype=np.uint8, mode="r")
            assert self._roll.size == 6 << 29
        return self._roll

    def gap_opt(self):
        g = _abi.GapOpt()
        fx.host_lib().fqb_gap_opt_default(C.byref(g))
        o = OrcGapOpt()
        for n, _ in OrcGapOpt._fields_:
            setattr(o, n, getattr(g, n))
        return o, g

    def align_batch(self, arrs, trim_qual=0, kmer_thresh=3, cap=8):
        """Restates bwa_read_seq_with_hash_dev's per-read prep + bwa_cal_sa_reg_gap for r = 2*pair+end."""
        L = self.lib
        n, rl = arrs[0].shape
        o, g = self.gap_opt()
        lens = np.zeros(2 * n, np.int32); filt = np.zeros(2 * n, np.uint8)
        out = np.zeros((2 * n, cap), _abi.ALN_DTYPE); na = np.zeros(2 * n, np.int32)
        max_len_diff = L.orc_cal_maxdiff(rl, C.c_double(0.02), C.c_double(g.fnr))
        slice_gapo = min(g.max_gapo, max_len_diff)
        n_buckets = (max_len_diff + 1) * g.s_mm + (slice_gapo + 1) * g.s_gapo + (g.max_gape + 1) * g.s_gape
        stack = C.c_void_p(L.orc_stack_new(n_buckets))
        roll = self.roll_tables() if kmer_thresh else None
        if roll is not None:
            base = roll.ctypes.data
            tabs = (C.c_void_p * 6)(*[base + (i << 29) for i in range(6)])
        for e in (0, 1):
            codes = fx.NT4[arrs[2 * e]]
            quals = arrs[2 * e + 1]
            for p in range(n):
                r = 2 * p + e
                c = np.ascontiguousarray(codes[p]); q = np.ascontiguousarray(quals[p])
                ln = L.orc_trim_len(trim_qual, _abi.u8p(q), rl)
                lens[r] = ln
                if kmer_thresh and not L.orc_kmer_pass(tabs, _abi.u8p(c), kmer_thresh):
                    filt[r] = 1
                    continue
                na[r] = L.orc_align_read(self.bwts, _abi.u8p(c), ln, C.byref(o), C.c_double(g.fnr), slice_gapo, stack,
                                         out[r].ctypes.data_as(C.c_void_p), cap)
        L.orc_stack_free(stack)
        return lens, filt, out, na
