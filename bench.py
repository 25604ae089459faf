#!/usr/bin/env python3
"""Contract benchmark: read-pairs/s of the FASTQuick align+summarize hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (torchrun launches it for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own CPU align on host cores)

A "step" is one pass of the hot path over one reference batch of 262,144 synthetic read
pairs (READ_BUFFER_SIZE, src/BwtMapper.h:37) drawn from BASELINE.json's configs[1]
workload (10M 2x100 bp pairs vs the 10k-marker index).  Every step uses a different batch;
reads shard across ranks with the index replicated (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from fastquick_b200 import _abi  # noqa: E402

BATCH = _abi.FQB_BATCH_PAIRS
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "FASTQuick_ref")
CLI_BIN = os.path.join(ROOT, "fastquick_b200", "FASTQuick_b200")
CLI_SAMPLE_BATCHES = 6           # batches of the workload the command-line leg runs (FASTQ text on disk: 2 x 220 B per pair)
NCU_DRAM_BYTES_PER_LAUNCH = 7.315e9   # search_kernel, one 262,144-pair launch of 2x100_10k: 3.476 GB read + 3.840 GB written (ncu capture r2a)
REF_SAMPLE_PAIRS = 32768         # pairs per step of the reference arm / cpu_baseline sample unit
STAGES = ("prep + k-mer filter + cal_width + match_gap + aln2seq/bwt_sa/mapQ + infer_isize + pairing + mate-rescue SW + gapped refinement + "
          "StatCollector pair classification and per-base pile-up/depth/quality/cycle accumulation (SURVEY 8 rows a1-a13)")
# BASELINE.json configs[1..4] (+ SURVEY 8(d)'s secondary WGS-like mix); the default and the contract line is configs[1].
# markers = (long, short, X, Y) flank counts of the synthetic reduced reference; reads = fqb_synth_read_cfg_t overrides
CONFIGS = {
    "2x100_10k": dict(read_len=100, markers=(1000, 9000, 100, 97), reads={},
                      workload="synthetic 10M 2x100bp pairs (f_on=1, 1% subst, 0.1%+0.1% indel) vs synthetic 10,197-marker flank index"),
    "2x150_10k": dict(read_len=150, markers=(1000, 9000, 100, 97), reads={},
                      workload="synthetic 200M 2x150bp pairs (f_on=1, 1% subst, 0.1%+0.1% indel) vs synthetic 10,197-marker flank index (BASELINE configs[2])"),
    "hapmap_higherr": dict(read_len=100, markers=(979, 8808, 0, 0), reads=dict(sub_rate=0.04, ins_rate=0.01, del_rate=0.01, max_indel_len=3),
                           workload="synthetic 2x100bp pairs at 4% subst, 1%+1% indel (lengths 1-3) vs synthetic 9,787-marker flank index (hapmap_3.3 count, BASELINE configs[3])"),
    "exome_2x150": dict(read_len=150, markers=(990, 8906, 0, 0), reads={},
                        workload="synthetic 2x150bp pairs vs synthetic 9,896-marker flank index (exome marker count, BASELINE configs[4])"),
    "wgs_mix": dict(read_len=100, markers=(1000, 9000, 100, 97), reads=dict(f_on=0.0021),
                    workload="synthetic 2x100bp pairs, WGS-like mix: 0.21% on target, the rest uniform background (filter-bound; SURVEY 8(d) secondary)"),
}
CFG = CONFIGS["2x100_10k"]
READ_LEN = 100
WORKLOAD = CFG["workload"]


def select_config(name):
    global CFG, READ_LEN, WORKLOAD
    CFG = CONFIGS[name]
    READ_LEN, WORKLOAD = CFG["read_len"], CFG["workload"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0, set()
        for t, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def make_synth(lib):
    cfg = _abi.SynthRefCfg()
    lib.fqb_synth_ref_cfg_default(C.byref(cfg))          # 1000 long + 9000 short + 100 X + 97 Y markers
    cfg.n_long, cfg.n_short, cfg.n_x, cfg.n_y = CFG["markers"]
    s = C.c_void_p()
    assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0, lib.fqb_last_error()
    return s


def gen_reads(lib, synth, first_pair, n_pairs, out=None):
    rc = _abi.SynthReadCfg()
    lib.fqb_synth_read_cfg_default(C.byref(rc))
    rc.read_len = READ_LEN
    for k, v in CFG["reads"].items():
        setattr(rc, k, v)
    arrs = out if out is not None else [np.empty((n_pairs, READ_LEN), np.uint8) for _ in range(4)]
    assert lib.fqb_synth_reads(synth, C.byref(rc), C.c_int64(first_pair), C.c_int64(n_pairs),
                               *[_abi.u8p(a) for a in arrs], 0) == 0, lib.fqb_last_error()
    return arrs


# ----------------------------------------------------------------------------- reference arm
def run_reference_sample(lib, synth, workdir, first_pair, n_pairs, index_prefix=None):
    """FASTQuick_ref align (the reference's own multithreaded CPU path) on n_pairs; returns (pairs/s, seconds, cores)."""
    if index_prefix is None:
        index_prefix = os.path.join(workdir, "bench.FASTQuick.fa")
        if not os.path.exists(index_prefix + ".rollhash"):
            assert lib.fqb_synth_write_inputs(synth, workdir.encode()) == 0, lib.fqb_last_error()
            assert lib.fqb_synth_write_index(synth, os.path.join(workdir, "genome.fa").encode(),
                                             os.path.join(workdir, "dbsnp.vcf").encode(), index_prefix.encode(), 1) == 0, lib.fqb_last_error()
    arrs = gen_reads(lib, synth, first_pair, n_pairs)
    fq = []
    for e in (0, 1):
        p = os.path.join(workdir, "s%d_%d.fq.gz" % (first_pair, e + 1))
        assert lib.fqb_write_fastq_gz(p.encode(), e + 1, C.c_int64(first_pair), C.c_int64(n_pairs), READ_LEN,
                                      _abi.u8p(arrs[2 * e]), _abi.u8p(arrs[2 * e + 1])) == 0
        fq.append(p)
    cores = os.cpu_count() or 1
    out_prefix = os.path.join(workdir, "out%d" % first_pair)
    cmd = [REF_BIN, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", index_prefix[: -len(".FASTQuick.fa")],
           "--out_prefix", out_prefix, "--t", str(cores), "--q", "15"]
    r = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    m = re.search(r"Processed Pair End mapping in ([0-9.]+) sec", r.stdout)
    if not m:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:])
    sec = float(m.group(1))
    for f in fq + [out_prefix + ".bam"]:
        try:
            os.remove(f)
        except OSError:
            pass
    return n_pairs / sec, sec, cores


def write_fastq_text(path, end, first_pair, bases, quals):
    """Plain-text FASTQ of one end (the records fqb_write_fastq_gz writes: @r%011d/<end>, bases, +, qualities), built with numpy."""
    n, L = bases.shape
    head = 16                                              # "@r" + 11 digits + "/e\n"
    rec = np.empty((n, head + L + 3 + L + 1), np.uint8)
    rec[:, 0], rec[:, 1] = ord("@"), ord("r")
    idx = np.arange(first_pair, first_pair + n, dtype=np.int64)
    for k in range(11):
        rec[:, 12 - k] = (idx // 10 ** k) % 10 + 48
    rec[:, 13], rec[:, 14], rec[:, 15] = ord("/"), 48 + end, 10
    rec[:, head:head + L] = bases
    rec[:, head + L:head + L + 3] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, head + L + 3:head + 2 * L + 3] = quals
    rec[:, -1] = 10
    with open(path, "wb") as f:
        f.write(rec.tobytes())


def feeder_rate(lib, fq, n_pairs):
    """The host feeder alone (fqb_feeder_fill: read + parse + SoA batches, one feeder per end side by side as in PairEndMapper)
    over the two FASTQ files; pairs/s, best of two passes."""
    lib.fqb_feeder_fill.restype = C.c_int64
    lib.fqb_feeder_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.fqb_feeder_close.argtypes = [C.c_void_p]
    cap = BATCH
    bufs = [[np.empty((cap, READ_LEN), np.uint8), np.empty((cap, READ_LEN), np.uint8), np.empty(cap, np.int32), np.empty((cap, 64), np.uint8)] for _ in (0, 1)]
    best = 0.0
    for _ in range(2):
        fd = []
        for p in fq:
            f = C.c_void_p()
            assert lib.fqb_feeder_open(p.encode(), 0, C.byref(f)) == 0, lib.fqb_last_error()
            fd.append(f)
        tot = [0, 0]

        def drain(e):
            b = bufs[e]
            while True:
                k = lib.fqb_feeder_fill(fd[e], cap, READ_LEN, b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, b[3].ctypes.data, 64)
                if k <= 0:
                    break
                tot[e] += k
        t0 = time.time()
        th = threading.Thread(target=drain, args=(0,))
        th.start(); drain(1); th.join()
        dt = time.time() - t0
        for f in fd:
            lib.fqb_feeder_close(f)
        if tot[0] == n_pairs and tot[1] == n_pairs:
            best = max(best, n_pairs / dt)
    return best


def run_cli_sample(lib, synth, workdir, n_pairs):
    """`FASTQuick_b200 align` -- the reference's command line on this library, FASTQ files in, statistics files and BAM out --
    on n_pairs of the workload as plain-text FASTQ; returns pairs/s by the CLI's own 'Processed Pair End mapping' line (what
    the reference arm reports for FASTQuick_ref) without and with BAM output."""
    index_prefix = os.path.join(workdir, "bench.FASTQuick.fa")
    if not os.path.exists(index_prefix + ".rollhash"):
        assert lib.fqb_synth_write_inputs(synth, workdir.encode()) == 0, lib.fqb_last_error()
        assert lib.fqb_synth_write_index(synth, os.path.join(workdir, "genome.fa").encode(),
                                         os.path.join(workdir, "dbsnp.vcf").encode(), index_prefix.encode(), 1) == 0, lib.fqb_last_error()
    fq = [os.path.join(workdir, "cli_%d.fq" % (e + 1)) for e in (0, 1)]
    files = [open(p, "wb") for p in fq]
    for f in files:
        f.close()
    chunk = BATCH
    for first in range(0, n_pairs, chunk):                 # the reads of bench steps 0, 1, ...: same generator, same seeds
        arrs = gen_reads(lib, synth, first, min(chunk, n_pairs - first))
        for e in (0, 1):
            part = fq[e] + ".part"
            write_fastq_text(part, e + 1, first, arrs[2 * e], arrs[2 * e + 1])
            with open(fq[e], "ab") as dst, open(part, "rb") as src:
                dst.write(src.read())
            os.remove(part)
    os.sync()                                              # the FASTQ text just written is on its way to disk: not during the timed runs
    out = {"feed": feeder_rate(lib, fq, n_pairs)}
    for tag, extra in (("stats", ["--sam_out"]), ("stats+bam", [])):
        cmd = [CLI_BIN, "align", "--fastq_1", fq[0], "--fastq_2", fq[1], "--index_prefix", index_prefix[: -len(".FASTQuick.fa")],
               "--out_prefix", os.path.join(workdir, "cli_out"), "--t", str(os.cpu_count() or 1), "--q", "15"] + extra
        r = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        m = re.search(r"Processed Pair End mapping in ([0-9.]+) sec", r.stdout)
        if r.returncode or not m:
            raise RuntimeError("FASTQuick_b200 align failed:\n" + r.stdout[-1500:])
        out[tag] = n_pairs / float(m.group(1))
    for f in fq + [os.path.join(workdir, "cli_out.bam")]:
        try:
            os.remove(f)
        except OSError:
            pass
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = _abi.load_library()
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/FASTQuick_ref not built (needs /root/reference at build time)"}))
        return
    synth = make_synth(lib)
    work = tempfile.mkdtemp(prefix="fqb_ref_")
    tot_pairs, tot_sec, cores = 0, 0.0, 0
    for step in range(args.warmup + args.steps):
        rate, sec, cores = run_reference_sample(lib, synth, work, step * REF_SAMPLE_PAIRS, REF_SAMPLE_PAIRS)
        if step >= args.warmup:
            tot_pairs += REF_SAMPLE_PAIRS; tot_sec += sec
    value = tot_pairs / tot_sec
    line = {
        "impl": "reference", "metric": "read-pairs/s (align+pileup)", "value": value, "unit": "read-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "name": args.config, "pairs_per_step": REF_SAMPLE_PAIRS,
                   "note": "reference FASTQuick align (k-mer filter, bwa aln/sampe, StatCollector, BAM) timed by its own "
                           "'Processed Pair End mapping' line; index load excluded"},
        "cpu_baseline": {"value": value, "unit": "read-pairs/s", "cores": cores, "kind": "reference",
                         "sample": "%d steps x %d pairs of the workload" % (args.steps, REF_SAMPLE_PAIRS)},
        "e2e": {"value": value, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
class Engine:
    """This rank's shard driven through the library's pipelined calls (interface of fastquick_b200.multigpu.run_sharded):
    batch b+world is submitted (upload + align stage on the align stream) before batch b is collected (pair / mate rescue /
    refinement / statistics on the main stream, with the hand-off ring on either side when world > 1)."""

    def __init__(self, lib, h, n_pairs, world):
        self.lib, self.h, self.n, self.world = lib, h, n_pairs, world
        self.src, self.on_device, self.rows = None, 0, None       # set per timed phase
        self.base = 0                                              # local step index of this phase's global batch 0
        self.packed_stride = 0                                     # > 0: the buffers hold the packed input form

    @staticmethod
    def _ptr(t):
        return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8))

    def submit(self, b):
        d = self.src[self.base + b // self.world]
        lib = self.lib
        if self.packed_stride:
            rc = lib.fqb_submit_pairs_packed(self.h, self.n, READ_LEN, self.packed_stride, self._ptr(d[0]), self._ptr(d[1]), None, self._ptr(d[2]), self._ptr(d[3]), None,
                                             self.on_device)
        else:
            rc = lib.fqb_submit_pairs(self.h, self.n, READ_LEN, self._ptr(d[0]), self._ptr(d[1]), None, self._ptr(d[2]), self._ptr(d[3]), None, self.on_device)
        assert rc == 0, lib.fqb_last_error()

    def collect(self, b, first_pair, is_last):
        lib = self.lib
        r1 = r2 = None
        if self.rows is not None:
            dst = self.rows[(b // self.world) & 1]
            r1, r2 = C.c_void_p(dst[0].data_ptr()), C.c_void_p(dst[1].data_ptr())
        assert lib.fqb_collect_pairs_sharded(self.h, r1, r2, C.c_uint64(b), C.c_uint64(first_pair), int(is_last)) == 0, lib.fqb_last_error()


def main_gpu(args):
    import torch
    import torch.distributed as dist
    from fastquick_b200 import multigpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _abi.load_library()      # raises when the CUDA extension is missing: no CPU fallback
    lib.fqb_stream.restype = C.c_void_p
    lib.fqb_launch_count.restype = C.c_uint64
    synth = make_synth(lib)
    g = _abi.GapOpt()
    lib.fqb_gap_opt_default(C.byref(g))
    g.trim_qual = 15                                  # bin/FASTQuick.sh --steps Align passes --q 15
    h = C.c_void_p()
    assert lib.fqb_create_from_synth(synth, C.byref(g), None, local, C.byref(h)) == 0, lib.fqb_last_error()
    # side files of the index (SelectedSite.vcf, .gc, dbSNP subset, .param, genome .fai/.amb) for the statistics tables
    work = tempfile.mkdtemp(prefix="fqb_bench_r%d_" % rank)
    prefix = os.path.join(work, "bench.FASTQuick.fa")
    assert lib.fqb_synth_write_inputs(synth, work.encode()) == 0, lib.fqb_last_error()
    assert lib.fqb_synth_write_index(synth, os.path.join(work, "genome.fa").encode(), os.path.join(work, "dbsnp.vcf").encode(), prefix.encode(), 0) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_open(h, prefix.encode()) == 0, lib.fqb_last_error()
    multigpu.bootstrap(lib, h, rank, world)          # NCCL communicator + hand-off mailboxes, inside the library
    stream = torch.cuda.ExternalStream(lib.fqb_stream(h))

    # BW_L2 of THIS box, measured now: random 64-byte reads over an 8 MiB buffer, all SMs, best of 10
    l2_best, l2_med = C.c_double(0), C.c_double(0)
    assert lib.fqb_measure_l2(local, C.c_int64(8 << 20), 64, 10, C.byref(l2_best), C.byref(l2_med)) == 0, lib.fqb_last_error()

    n_steps = args.warmup + args.steps
    n_pairs = args.pairs_per_step
    # this rank's shard of the workload: global batch b = s * world + rank
    # Input form: what fqb_feeder_fill_packed emits -- 2-bit bases (32 / 48 bytes per 100 / 150-base read) and the quality
    # bytes with the not-ACGT flag (--ascii-input: the ASCII rows of fqb_feeder_fill).  The batches are packed here, once,
    # outside every timed region (in the CLI the feeder's parse workers do it).
    lib.fqb_packed_stride.restype = C.c_int32
    pstride = 0 if args.ascii_input else lib.fqb_packed_stride(READ_LEN)
    host = []
    scratch = [np.empty((n_pairs, READ_LEN), np.uint8) for _ in range(4)] if pstride else None
    for s in range(n_steps):
        first = (s * world + rank) * n_pairs
        if pstride:
            gen_reads(lib, synth, first, n_pairs, out=scratch)
            bufs = [torch.empty((n_pairs, pstride if (i & 1) == 0 else READ_LEN), dtype=torch.uint8).pin_memory() for i in range(4)]
            for e in (0, 2):
                assert lib.fqb_pack_reads(C.c_int64(n_pairs), READ_LEN, C.c_void_p(scratch[e].ctypes.data), C.c_void_p(scratch[e + 1].ctypes.data), pstride,
                                          C.c_void_p(bufs[e].data_ptr()), C.c_void_p(bufs[e + 1].data_ptr())) == 0, lib.fqb_last_error()
        else:
            bufs = [torch.empty((n_pairs, READ_LEN), dtype=torch.uint8).pin_memory() for _ in range(4)]
            gen_reads(lib, synth, first, n_pairs, out=[b.numpy() for b in bufs])
        host.append(bufs)
    dev = [[b.cuda(non_blocking=True) for b in bufs] for bufs in host]   # whole shard resident in HBM
    torch.cuda.synchronize()
    eng = Engine(lib, h, n_pairs, world)
    eng.packed_stride = pstride
    h2d_bytes = sum(int(b.numel()) for b in host[0])
    rows_host = [[torch.empty((n_pairs, _abi.READ_DTYPE.itemsize), dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(2)]
    merge_ms = []

    def merge():
        ms = C.c_double(0)
        assert lib.fqb_comm_merge_stats(h, C.byref(ms)) == 0, lib.fqb_last_error()      # grouped ncclReduce + exact-size ncclSend/Recv
        merge_ms.append(ms.value)

    def run(first_step, k, e2e):
        # device arm: inputs resident in HBM.  e2e arm: the public calls on pinned host buffers -- every step uploads its
        # 4 input arrays and copies its result rows back, both inside the timed region
        eng.base = first_step
        eng.src, eng.on_device, eng.rows = (host, 0, rows_host) if e2e else (dev, 1, None)
        multigpu.run_sharded(eng, k * world, rank, world, n_pairs)
        assert lib.fqb_rows_wait(h) == 0, lib.fqb_last_error()       # the one wait of the run: rows of the last batch + deferred status

    def rq_time():
        ms, nl = C.c_double(0.0), C.c_uint64(0)
        assert lib.fqb_rank_query_time(h, C.byref(ms), C.byref(nl)) == 0
        return ms.value, int(nl.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e):
        lib.fqb_reset_stream(h)
        run(0, args.warmup, e2e)
        merge()                     # warm-up of the end-of-run exchange too (NCCL sets up its channels lazily)
        assert lib.fqb_stats_reset(h) == 0, lib.fqb_last_error()     # the timed region starts from empty accumulators
        barrier()
        lib.fqb_reset_stream(h)
        c0 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c0)
        l0 = lib.fqb_launch_count(h)
        rq0 = rq_time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        run(args.warmup, args.steps, e2e)
        merge()
        e1.record(stream)
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        c1 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c1)
        launches = lib.fqb_launch_count(h) - l0
        rq1 = rq_time()
        rq = (rq1[0] - rq0[0], rq1[1] - rq0[1])
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, [int(c1[i] - c0[i]) for i in range(3)], int(launches), t0, t1, rq

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # wait for the sampler's first line: on a fresh box nvidia-smi takes a second to initialise NVML, and that start-up holds
        # driver locks -- inside the timed region it showed up as a one-off 100 ms stall of the first arm
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 10.0:
            time.sleep(0.05)
        time.sleep(0.2)
    ms, ctr, launches, t0, t1, rq = timed(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    merge_ms_device = merge_ms[-1]
    ms_e2e = timed(True)[0]

    # the dominant kernels alone: in the pipelined loop the next batch's search shares the device with the later stages of the
    # current one (and with the draining tail of the previous search), so a CUDA-event bracket there also counts time spent
    # waiting for SM room.  For the roofline the same kernels are timed with nothing else on the device: the last batches of
    # the timed region once more, stage by stage on one stream (fqb_stage_align), events around width + order + search.
    iso_steps = min(3, args.steps)
    c0 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c0)
    rq0 = rq_time()
    for s_ in range(n_steps - iso_steps, n_steps):
        d = dev[s_]
        if pstride:
            rc = lib.fqb_stage_load_packed(h, n_pairs, READ_LEN, pstride, Engine._ptr(d[0]), Engine._ptr(d[1]), None, Engine._ptr(d[2]), Engine._ptr(d[3]), None, 1)
        else:
            rc = lib.fqb_stage_load(h, n_pairs, READ_LEN, Engine._ptr(d[0]), Engine._ptr(d[1]), None, Engine._ptr(d[2]), Engine._ptr(d[3]), None, 1)
        assert rc == 0, lib.fqb_last_error()
        assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()
    c1 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c1)
    rq1 = rq_time()
    iso_ms, iso_n, iso_blk = rq1[0] - rq0[0], rq1[1] - rq0[1], int(c1[2] - c0[2])

    total_pairs = args.steps * n_pairs * world
    value = total_pairs / (ms * 1e-3)
    e2e_value = total_pairs / (ms_e2e * 1e-3)
    # roofline of the dominant kernels (bwt_cal_width + bwt_match_gap): algorithmic bytes = 64 B x N_blk,
    # N_blk counted on the device for exactly the reads processed in the timed region; clock = CUDA events recorded by
    # the engine around those launches on the stream they run on, averaged over the launches of the timed region
    hbm_peak, peak_kind = peaks()
    n_blk = ctr[2]
    rq_ms_in_pipeline = rq[0] / max(rq[1], 1)
    rq_ms_per_launch = iso_ms / max(iso_n, 1)
    bytes_per_launch = 64.0 * iso_blk / max(iso_n, 1)
    achieved = bytes_per_launch / (rq_ms_per_launch * 1e-3) / 1e9
    touches_job = float(n_blk)
    if world > 1:
        t = torch.tensor([touches_job], device="cuda", dtype=torch.float64); dist.all_reduce(t); touches_job = float(t.item())
    touches_per_s_job = touches_job / (ms * 1e-3)
    if rank != 0:
        lib.fqb_destroy(h)
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": "read-pairs/s (align+pileup)", "value": value, "unit": "read-pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "name": args.config, "pairs_per_step": n_pairs, "read_len": READ_LEN, "stages": STAGES,
                   "l2_policy": "every step reads a different %d MB batch; per-step working set (inputs+widths+stack arena) exceeds the 126 MB L2" % (h2d_bytes // 1000000),
                   "input_form": ("ASCII bases + qualities (fqb_feeder_fill)" if not pstride else
                                  "packed: 2-bit bases, %d bytes per read, + quality bytes with the not-ACGT flag (what fqb_feeder_fill_packed emits; packed once, outside the timed regions)" % pstride),
                   "index": "%d markers (counts of the named marker set; positions/alleles synthetic), replicated per GPU" % sum(CFG["markers"]),
                   "pipeline": "fqb_submit_pairs(b+1) before fqb_collect_pairs_sharded(b): align stage of the next batch on a second stream, "
                               "no host synchronisation per batch (one wait at the end of the run)",
                   "multi_gpu": "batches round-robin over ranks; drand48 position + last_ii handed rank to rank through NVLink peer memory "
                                "(56 B per batch, one-thread kernels); accumulators reduced with one NCCL group and pile-up entries / "
                                "duplicate keys sent to rank 0 inside the timed region, all in the C library",
                   "reference_arm_sample": "bench.py --impl reference times %d-pair samples of the same workload per step (the CPU path is ~1000x slower)" % REF_SAMPLE_PAIRS},
        "e2e": {"value": e2e_value, "unit": "read-pairs/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(2 * n_pairs * _abi.READ_DTYPE.itemsize), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "l2", "achieved": achieved, "peak": l2_best.value, "unit": "GB/s", "frac": achieved / l2_best.value,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (n_pairs == BATCH and args.config == "2x100_10k") else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of search_kernel in its plain form (FQB_SEARCH_VAR=0), ncu --set full capture r2a "
                                       "(profiles/r02_search_kernel_ncu.md); the default form since -- pop staging + streaming stack stores, "
                                       "profiles/r02_search_memory_variants.md -- moves the same entries and was not re-captured",
                     "peak_kind": "BW_L2 measured on this box in this run: fqb_measure_l2, random 64-byte reads over an 8 MiB buffer, all SMs, best of 10 (median %.1f)" % l2_med.value,
                     "hbm_peak": hbm_peak, "hbm_peak_kind": peak_kind + " HBM copy bandwidth (MEASURED_PEAKS.json)", "frac_of_hbm_peak": achieved / hbm_peak,
                     "kernel": "width_kernel + search_kernel (rank queries of one %d-pair batch)" % n_pairs,
                     "kernel_ms_per_launch": rq_ms_per_launch, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "kernel_timing": "CUDA events around width + order + search on their stream, the last %d batches of the timed region run once more "
                                      "stage by stage with nothing else on the device" % iso_steps,
                     "kernel_ms_in_pipeline": rq_ms_in_pipeline,
                     "kernel_ms_in_pipeline_note": "the same bracket inside the timed region: includes waiting for SM room next to the other stream's kernels",
                     "share_of_step": rq_ms_per_launch / (ms / args.steps),
                     "algorithmic": "64 B x N_blk occ-block touches of bwt_cal_width+bwt_match_gap, per GPU; N_blk/pair = %.1f" % (n_blk / (args.steps * n_pairs)),
                     "note": "the FM index (10 MB) is L2-resident; the kernels are bound by dependent-latency x steps and instruction issue, not by bandwidth (DESIGN.md 3.3)",
                     "occ_block_touches_per_s_job": touches_per_s_job},
        "exchange_ms": merge_ms_device,
    }
    lib.fqb_destroy(h)                   # the host legs below run other processes on this GPU / these cores
    workc = None
    if world == 1 and not args.no_cli:
        # like for like with the reference arm (its timer covers FASTQ decoding, the statistics text and the BAM): the same
        # command line, on this library, timed by the same log line
        try:
            workc = tempfile.mkdtemp(prefix="fqb_cpu_")
            n_c = CLI_SAMPLE_BATCHES * BATCH
            rates = run_cli_sample(lib, synth, workc, n_c)
            line["cli"] = {"value": rates["stats+bam"], "unit": "read-pairs/s", "value_without_bam": rates["stats"], "pairs": n_c,
                           "feeder_alone": rates["feed"],
                           "what": "FASTQuick_b200 align --device 0 on %d pairs of the workload as plain-text FASTQ (feeder, upload, all stages, "
                                   "InsertSizeTable text, BAM), by its 'Processed Pair End mapping' line; index load and the final "
                                   "ProcessCore files excluded, as in the reference arm; feeder_alone = fqb_feeder_fill over the same two files, pairs/s" % n_c}
        except Exception as ex:
            line["cli"] = {"value": None, "unit": "read-pairs/s", "what": "failed: %s" % str(ex)[:300]}
    if world == 1 and not args.no_cpu_baseline:
        try:
            workc = workc or tempfile.mkdtemp(prefix="fqb_cpu_")
            n_s = REF_SAMPLE_PAIRS
            if os.path.exists(REF_BIN):
                rate, sec, cores = run_reference_sample(lib, synth, workc, 0, n_s)
                line["cpu_baseline"] = {"value": rate, "unit": "read-pairs/s", "cores": cores, "kind": "reference",
                                        "sample": "first %d pairs of the workload through FASTQuick_ref align --t %d --q 15 (%.1f s mapping)" % (n_s, cores, sec)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "read-pairs/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref/FASTQuick_ref missing"}
        except Exception as ex:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": "read-pairs/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % str(ex)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2x100_10k", choices=sorted(CONFIGS), help="workload (default: BASELINE.json configs[1], the contract line)")
    ap.add_argument("--pairs-per-step", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the FASTQuick_b200 command-line leg (FASTQ files in, statistics files / BAM out)")
    ap.add_argument("--ascii-input", action="store_true", help="hand the batches over as ASCII rows instead of the packed form")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
