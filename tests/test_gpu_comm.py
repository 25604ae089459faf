"""Row (e), native: the hand-off ring and the end-of-run merge inside the C library.

test_ring_two_handles_one_gpu -- two handles of one process (fqb_comm_init_local: mailboxes wired directly) deal the batches
of a file between them through fqb_submit_pairs / fqb_collect_pairs_sharded; rows of every batch must equal a single handle's.
test_two_ranks_nccl_equals_one_rank -- needs two GPUs: two PROCESSES, one per GPU (mailboxes through CUDA IPC, NCCL communicator
from fqb_comm_init), run the sharded loop, fqb_comm_merge_stats, and rank 0 writes the files: every output file must equal
the single-rank run's."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import fx
from fastquick_b200 import _abi
from test_gpu_sharded import STAT_FILES, _handle, _batch

pytestmark = pytest.mark.gpu
BATCH, N_BATCHES = 2500, 6


def _subs(index):
    arrs = index.reads(N_BATCHES * BATCH, read_len=100, seed=4242, ins_rate=0.003, del_rate=0.003)
    return [[np.ascontiguousarray(a[b * BATCH:(b + 1) * BATCH]) for a in arrs] for b in range(N_BATCHES)]


def _single_run(lib, index, subs, prefix):
    h = _handle(index, prefix)
    rows = [_batch(lib, h, subs[b], b * BATCH, None)[1] for b in range(N_BATCHES)]
    assert lib.fqb_stats_finish(h, prefix.encode()) == 0, lib.fqb_last_error()
    lib.fqb_destroy(h)
    return rows


def _submit(lib, h, s):
    assert lib.fqb_submit_pairs(h, BATCH, 100, _abi.u8p(s[0]), _abi.u8p(s[1]), None, _abi.u8p(s[2]), _abi.u8p(s[3]), None, 0) == 0, lib.fqb_last_error()


def test_ring_two_handles_one_gpu(small_index, tmp_path):
    lib = fx.host_lib()
    subs = _subs(small_index)
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        rows_one = _single_run(lib, small_index, subs, str(tmp_path / "one"))
        hs = [_handle(small_index, str(tmp_path / ("ring%d" % r))) for r in range(2)]
        arr = (C.c_void_p * 2)(hs[0], hs[1])
        assert lib.fqb_comm_init_local(arr, 2) == 0, lib.fqb_last_error()
        rows = [[np.zeros(BATCH, _abi.READ_DTYPE) for _ in range(2)] for _ in range(N_BATCHES)]
        _submit(lib, hs[0], subs[0]); _submit(lib, hs[1], subs[1])
        for b in range(N_BATCHES):
            h = hs[b % 2]
            assert lib.fqb_collect_pairs_sharded(h, rows[b][0].ctypes.data_as(C.c_void_p), rows[b][1].ctypes.data_as(C.c_void_p),
                                                 C.c_uint64(b), C.c_uint64(b * BATCH), int(b == N_BATCHES - 1)) == 0, lib.fqb_last_error()
            if b + 2 < N_BATCHES:
                _submit(lib, h, subs[b + 2])
        for h in hs:
            assert lib.fqb_rows_wait(h) == 0, lib.fqb_last_error()
            lib.fqb_destroy(h)
    finally:
        os.chdir(cwd)
    for b in range(N_BATCHES):
        for e in (0, 1):
            assert rows[b][e].tobytes() == rows_one[b][e].tobytes(), "batch %d end %d" % (b, e)


def _rank_worker(rank, world, index_dir, out_dir, port):
    """One process per GPU: the sharded loop of fastquick_b200.multigpu with the real engine."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fastquick_b200 import multigpu
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _abi.load_library()
    index = fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)
    subs = _subs(index)
    os.chdir(index.dir)
    prefix = os.path.join(out_dir, "r%d" % rank)
    g = _abi.GapOpt(); lib.fqb_gap_opt_default(C.byref(g)); g.trim_qual = 15
    h = C.c_void_p()
    assert lib.fqb_create(index.prefix.encode(), C.byref(g), None, rank, C.byref(h)) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_open(h, index.prefix.encode()) == 0, lib.fqb_last_error()
    assert lib.fqb_stats_begin_file(h, prefix.encode(), b"r1.fq", b"r2.fq") == 0, lib.fqb_last_error()
    multigpu.bootstrap(lib, h, rank, world)
    rows = {}

    class Eng:
        def submit(self, b):
            _submit(lib, h, subs[b])

        def collect(self, b, first_pair, is_last):
            rows[b] = [np.zeros(BATCH, _abi.READ_DTYPE) for _ in range(2)]
            assert lib.fqb_collect_pairs_sharded(h, rows[b][0].ctypes.data_as(C.c_void_p), rows[b][1].ctypes.data_as(C.c_void_p),
                                                 C.c_uint64(b), C.c_uint64(first_pair), int(is_last)) == 0, lib.fqb_last_error()
            assert lib.fqb_stats_emit(h, None, 0) == 0, lib.fqb_last_error()

    multigpu.run_sharded(Eng(), N_BATCHES, rank, world, BATCH)
    assert lib.fqb_rows_wait(h) == 0, lib.fqb_last_error()
    ms = C.c_double(0)
    assert lib.fqb_comm_merge_stats(h, C.byref(ms)) == 0, lib.fqb_last_error()
    if rank != 0:
        assert lib.fqb_stats_close_table(h) == 0, lib.fqb_last_error()
    dist.barrier()
    if rank == 0:
        others = (C.c_char_p * (world - 1))(*[os.path.join(out_dir, "r%d" % r).encode() for r in range(1, world)])
        assert lib.fqb_stats_merge_tables(h, others, world - 1) == 0, lib.fqb_last_error()
        assert lib.fqb_stats_finish(h, prefix.encode()) == 0, lib.fqb_last_error()
    np.savez(os.path.join(out_dir, "rows%d.npz" % rank), **{"b%d_%d" % (b, e): rows[b][e] for b in rows for e in (0, 1)})
    lib.fqb_destroy(h)
    dist.destroy_process_group()


def test_two_ranks_nccl_equals_one_rank(small_index, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    lib = fx.host_lib()
    subs = _subs(small_index)
    cwd = os.getcwd()
    os.chdir(small_index.dir)
    try:
        one = str(tmp_path / "one")
        rows_one = _single_run(lib, small_index, subs, one)
    finally:
        os.chdir(cwd)
    out_dir = str(tmp_path / "sharded")
    os.makedirs(out_dir)
    mp.spawn(_rank_worker, args=(2, small_index.dir, out_dir, 29611), nprocs=2, join=True)
    for r in range(2):
        z = np.load(os.path.join(out_dir, "rows%d.npz" % r))
        for b in range(r, N_BATCHES, 2):
            for e in (0, 1):
                assert z["b%d_%d" % (b, e)].tobytes() == rows_one[b][e].tobytes(), "batch %d end %d" % (b, e)
    two = os.path.join(out_dir, "r0")
    for ext in STAT_FILES:
        a = [l for l in open(one + "." + ext) if not l.startswith("##fileDate")]
        b = [l for l in open(two + "." + ext) if not l.startswith("##fileDate")]
        assert a == b, ext


def test_cli_two_devices_equal_one_device(small_index, tmp_path):
    """`FASTQuick_b200 align --devices 0,1` (C++ BwtMapper dealing the batches over two engines, hand-off ring and NCCL merge
    inside the library) against `--device 0`: every statistics file and every BAM record.  The batch size is the reference's
    262,144 pairs, so the input is two FASTQ pairs of 300,000 pairs each through --fq_list: two batches per file, and the
    ring starts a new epoch (srand48, last_ii) with the second file."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import subprocess
    from test_gpu_stats import TEXT_FILES, _compare_files
    from test_gpu_cli import CLI, _compare_bams
    lines = []
    for k in (0, 1):
        arrs = small_index.reads(300000, read_len=100, seed=2024 + k, ins_rate=0.002, del_rate=0.002, first_pair=k * 300000)
        fq = small_index.write_fastq("clidev%d" % k, arrs, first_pair=k * 300000)
        lines.append(fq[0] + "\t" + fq[1])
    fq_list = os.path.join(small_index.dir, "clidev.list")
    open(fq_list, "w").write("\n".join(lines) + "\n")
    idx_prefix = small_index.prefix[: -len(".FASTQuick.fa")]
    outs = {}
    for tag, dev in (("one", ["--device", "0"]), ("two", ["--devices", "0,1"])):
        out = os.path.join(small_index.dir, "clidev_" + tag)
        cmd = [CLI, "align", "--fq_list", fq_list, "--index_prefix", idx_prefix, "--out_prefix", out, "--t", "8", "--q", "15"] + dev
        r = subprocess.run(cmd, cwd=small_index.dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
        outs[tag] = out
    for ext in TEXT_FILES + ["FASTQ.csv"]:
        _compare_files(outs["one"] + "." + ext, outs["two"] + "." + ext)
    va = [l for l in open(outs["one"] + ".vcf") if not l.startswith("##fileDate")]
    vb = [l for l in open(outs["two"] + ".vcf") if not l.startswith("##fileDate")]
    assert va == vb
    recs = _compare_bams(outs["one"] + ".bam", outs["two"] + ".bam")
    assert len(recs) > 1000000
    assert not os.path.exists(outs["two"] + ".shard1.InsertSizeTable")
