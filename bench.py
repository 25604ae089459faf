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
READ_LEN = 100
WORKLOAD = "synthetic 10M 2x100bp pairs (f_on=1, 1% subst, 0.1%+0.1% indel) vs synthetic 10,197-marker flank index"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "FASTQuick_ref")
NCU_DRAM_BYTES_PER_LAUNCH = 7.341e9   # search_kernel, one 262,144-pair launch: 3.499 GB read + 3.843 GB written (ncu capture r1f)
REF_SAMPLE_PAIRS = 32768         # pairs per step of the reference arm / cpu_baseline sample unit
STAGES = ("prep + k-mer filter + cal_width + match_gap + aln2seq/bwt_sa/mapQ + infer_isize + pairing + mate-rescue SW + gapped refinement + "
          "StatCollector pair classification and per-base pile-up/depth/quality/cycle accumulation (SURVEY 8 rows a1-a13)")


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
    s = C.c_void_p()
    assert lib.fqb_synth_create(C.byref(cfg), C.byref(s)) == 0, lib.fqb_last_error()
    return s


def gen_reads(lib, synth, first_pair, n_pairs, out=None):
    rc = _abi.SynthReadCfg()
    lib.fqb_synth_read_cfg_default(C.byref(rc))
    rc.read_len = READ_LEN
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
        "config": {"workload": WORKLOAD, "pairs_per_step": REF_SAMPLE_PAIRS,
                   "note": "reference FASTQuick align (k-mer filter, bwa aln/sampe, StatCollector, BAM) timed by its own "
                           "'Processed Pair End mapping' line; index load excluded"},
        "cpu_baseline": {"value": value, "unit": "read-pairs/s", "cores": cores, "kind": "reference",
                         "sample": "%d steps x %d pairs of the workload" % (args.steps, REF_SAMPLE_PAIRS)},
        "e2e": {"value": value, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
class Engine:
    """The per-batch stages of the hot path over this rank's resident shard (interface of fastquick_b200.multigpu)."""

    def __init__(self, lib, h, dev, n_pairs, rank, world):
        self.lib, self.h, self.dev, self.n, self.rank, self.world = lib, h, dev, n_pairs, rank, world
        self.base = 0           # local step index of global batch 0 of the current phase

    def _ptr(self, t):
        return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8))

    def align(self, b):
        d = self.dev[self.base + b // self.world]
        lib, h = self.lib, self.h
        assert lib.fqb_stage_load(h, self.n, READ_LEN, self._ptr(d[0]), self._ptr(d[1]), None, self._ptr(d[2]), self._ptr(d[3]), None, 1) == 0, lib.fqb_last_error()
        assert lib.fqb_set_pair_base(h, C.c_uint64(b * self.n)) == 0
        assert lib.fqb_stage_align(h) == 0, lib.fqb_last_error()

    def pair(self, b):
        assert self.lib.fqb_stage_pair(self.h) == 0, self.lib.fqb_last_error()

    def finish(self, b):
        assert self.lib.fqb_stage_sw_refine(self.h) == 0, self.lib.fqb_last_error()
        assert self.lib.fqb_stage_stats(self.h) == 0, self.lib.fqb_last_error()

    def var_export(self, which):
        import torch
        n = C.c_uint64(0)
        assert self.lib.fqb_stats_var_count(self.h, which, C.byref(n)) == 0, self.lib.fqb_last_error()
        item = 20 if which == 0 else 8
        t = torch.empty(int(n.value) * item, dtype=torch.uint8, device="cuda")       # exported device to device
        if n.value:
            assert self.lib.fqb_stats_var_export(self.h, which, C.c_void_p(t.data_ptr()), n) == 0, self.lib.fqb_last_error()
        return t

    def var_import(self, which, t):
        item = 20 if which == 0 else 8
        t = t.contiguous()
        assert self.lib.fqb_stats_var_import(self.h, which, C.c_void_p(t.data_ptr()), C.c_uint64(t.numel() // item)) == 0, self.lib.fqb_last_error()

    def get_state(self):
        calls, ii = C.c_uint64(0), _abi.ISize()
        self.lib.fqb_get_stream_state(self.h, C.byref(calls), C.byref(ii))
        raw = np.frombuffer(bytes(ii), dtype=np.int64)
        return [int(calls.value)] + [int(x) for x in raw] + [0] * (7 - len(raw))

    def set_state(self, s):
        ii = _abi.ISize.from_buffer_copy(np.array(s[1:1 + C.sizeof(_abi.ISize) // 8], dtype=np.int64).tobytes())
        self.lib.fqb_set_stream_state(self.h, C.c_uint64(int(s[0])), C.byref(ii))


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
    stream = torch.cuda.ExternalStream(lib.fqb_stream(h))

    n_steps = args.warmup + args.steps
    n_pairs = args.pairs_per_step
    # this rank's shard of the workload: global batch b = s * world + rank
    host = []
    for s in range(n_steps):
        first = (s * world + rank) * n_pairs
        bufs = [torch.empty((n_pairs, READ_LEN), dtype=torch.uint8).pin_memory() for _ in range(4)]
        gen_reads(lib, synth, first, n_pairs, out=[b.numpy() for b in bufs])
        host.append(bufs)
    dev = [[b.cuda(non_blocking=True) for b in bufs] for bufs in host]   # whole shard resident in HBM
    torch.cuda.synchronize()
    eng = Engine(lib, h, dev, n_pairs, rank, world)

    # accumulator groups reduced to rank 0 at the end of the timed region (NCCL over NVLink)
    groups = []
    for which, dt, op in ((0, torch.int32, "sum"), (1, torch.int64, "sum"), (2, torch.int32, "sum"), (3, torch.int32, "min")):
        nb = C.c_uint64(0)
        assert lib.fqb_stats_group_bytes(h, which, C.byref(nb)) == 0
        groups.append((which, torch.empty(int(nb.value) // (4 if dt == torch.int32 else 8), dtype=dt, device=device), op))

    def reduce_stats():
        if world == 1:
            return
        tt = [time.time()]
        for which, t, op in groups:
            assert lib.fqb_stats_export(h, which, C.c_void_p(t.data_ptr())) == 0, lib.fqb_last_error()
        multigpu.reduce_accumulators([(t, op) for _, t, op in groups], rank, world)
        if rank == 0:
            for which, t, op in groups:
                assert lib.fqb_stats_import(h, which, C.c_void_p(t.data_ptr())) == 0, lib.fqb_last_error()
        torch.cuda.synchronize(); tt.append(time.time())
        multigpu.gather_variable(eng, rank, world, device)      # pile-up entries + duplicate keys, after the sums
        torch.cuda.synchronize(); tt.append(time.time())
        if os.environ.get("FQB_BENCH_VERBOSE") and rank == 0:
            print("reduce: fixed groups %.1f ms, variable state %.1f ms" % ((tt[1] - tt[0]) * 1e3, (tt[2] - tt[1]) * 1e3), file=sys.stderr, flush=True)

    def ptr(t):
        return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_uint8))

    rows_host = [[torch.empty((n_pairs, _abi.READ_DTYPE.itemsize), dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(2)]
    ii_host = _abi.ISize()

    def run_device(first_step, k):
        eng.base = first_step
        multigpu.run_sharded(eng, k * world, rank, world, device)

    def run_e2e(first_step, k):
        # the public per-batch calls: pinned host FASTQ arrays in, per-read result rows out, statistics accumulated.
        # Uploads of batch s+1 and the row copies of batch s overlap the kernels of the neighbouring batches.
        for s in range(first_step, first_step + k):
            b = host[s]
            if s + 1 < first_step + k:
                nb = host[s + 1]
                assert lib.fqb_prefetch_pairs(h, n_pairs, READ_LEN, ptr(nb[0]), ptr(nb[1]), None, ptr(nb[2]), ptr(nb[3]), None) == 0, lib.fqb_last_error()
            rc = lib.fqb_align_pairs(h, n_pairs, READ_LEN, ptr(b[0]), ptr(b[1]), None, ptr(b[2]), ptr(b[3]), None, None, None, C.byref(ii_host))
            assert rc == 0, lib.fqb_last_error()
            assert lib.fqb_stage_stats(h) == 0, lib.fqb_last_error()
            dst = rows_host[s & 1]
            assert lib.fqb_stage_fetch_rows_async(h, C.c_void_p(dst[0].data_ptr()), C.c_void_p(dst[1].data_ptr())) == 0, lib.fqb_last_error()
        assert lib.fqb_rows_wait(h) == 0, lib.fqb_last_error()

    def rq_time():
        ms, nl = C.c_double(0.0), C.c_uint64(0)
        assert lib.fqb_rank_query_time(h, C.byref(ms), C.byref(nl)) == 0
        return ms.value, int(nl.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, with_reduce):
        lib.fqb_reset_stream(h)
        fn(0, args.warmup)
        if with_reduce:
            reduce_stats()          # warm-up of the end-of-run exchange too (NCCL sets up its channels lazily)
        barrier()
        lib.fqb_reset_stream(h)
        c0 = (C.c_uint64 * 4)(); lib.fqb_stage_counters(h, c0)
        l0 = lib.fqb_launch_count(h)
        rq0 = rq_time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn(args.warmup, args.steps)
            if with_reduce:
                reduce_stats()
                stream.wait_stream(torch.cuda.current_stream())
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
        time.sleep(0.3)
    ms, ctr, launches, t0, t1, rq = timed(run_device, True)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_e2e = timed(run_e2e, True)[0]

    total_pairs = args.steps * n_pairs * world
    value = total_pairs / (ms * 1e-3)
    e2e_value = total_pairs / (ms_e2e * 1e-3)
    # roofline of the dominant kernels (bwt_cal_width + bwt_match_gap): algorithmic bytes = 64 B x N_blk,
    # N_blk counted on the device for exactly the reads processed in the timed region; clock = CUDA events recorded by
    # the engine around those launches on its own stream, averaged over the launches of the timed region
    hbm_peak, peak_kind = peaks()
    n_blk = ctr[2]
    rq_ms_per_launch = rq[0] / max(rq[1], 1)
    bytes_per_launch = 64.0 * n_blk / max(rq[1], 1)
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
        "config": {"workload": WORKLOAD, "pairs_per_step": n_pairs, "read_len": READ_LEN, "stages": STAGES,
                   "l2_policy": "every step reads a different 105 MB batch; per-step working set (inputs+widths+stack arena) exceeds the 126 MB L2",
                   "index": "10,197 markers, l_pac 6,608,697, replicated per GPU",
                   "multi_gpu": "batches round-robin over ranks; drand48 position + last_ii handed rank to rank (56 B per batch); "
                                "accumulators NCCL-reduced and pile-up entries / duplicate keys gathered to rank 0 inside the timed region"},
        "e2e": {"value": e2e_value, "unit": "read-pairs/s", "h2d_bytes_per_step": 4 * n_pairs * READ_LEN,
                "d2h_bytes_per_step": int(2 * n_pairs * _abi.READ_DTYPE.itemsize)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if n_pairs == BATCH else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of search_kernel, ncu --set full capture r1f (profiles/r01_search_kernel_ncu.md)",
                     "peak_kind": peak_kind + " HBM copy bandwidth (MEASURED_PEAKS.json)",
                     "kernel": "width_kernel + search_kernel (rank queries of one 262,144-pair batch)",
                     "kernel_ms_per_launch": rq_ms_per_launch, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "share_of_step": rq_ms_per_launch / (ms / args.steps),
                     "algorithmic": "64 B x N_blk occ-block touches of bwt_cal_width+bwt_match_gap, per GPU; N_blk/pair = %.1f" % (n_blk / (args.steps * n_pairs)),
                     "note": "the FM index (10 MB) is L2-resident: these kernels are bound by issue slots and stack-pop latency, not by HBM bandwidth (DESIGN.md 3.3)",
                     "occ_block_touches_per_s_job": touches_per_s_job},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            workc = tempfile.mkdtemp(prefix="fqb_cpu_")
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
    lib.fqb_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
