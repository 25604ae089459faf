"""world_size-2 gloo test of the multi-GPU protocol (fastquick_b200/multigpu.py) with a stand-in engine: the drand48
position and last_ii handed along the ring must reproduce exactly what a single process sees, and the accumulator
reduction must equal the single-process totals."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from fastquick_b200 import multigpu  # noqa: E402


class FakeEngine:
    """Consumes a batch-dependent number of 'draws'; a batch with no confident pairs reuses the previous estimate."""

    def __init__(self):
        self.calls, self.ii, self.log, self.acc = 0, [0] * 7, {}, torch.zeros(16, dtype=torch.int64)
        self.var, self.got = {0: [], 1: []}, {}

    def align(self, b):
        pass

    def pair(self, b):
        self.log[b] = (self.calls, tuple(self.ii))       # state this batch started from
        self.calls += 1000 + 17 * b
        if b % 3 != 2:                                    # inference "succeeds"
            self.ii = [b * 11 + k for k in range(7)]

    def finish(self, b):
        self.acc[b % 16] += b + 1
        self.var[0] += [b] * (b % 3)          # a batch-dependent amount of variable-size state
        self.var[1] += [100 + b]

    def var_export(self, which):
        return torch.tensor(self.var[which], dtype=torch.uint8)

    def var_import(self, which, t):
        self.got.setdefault(which, []).append(t.tolist())

    def get_state(self):
        return [self.calls] + list(self.ii)

    def set_state(self, s):
        self.calls, self.ii = int(s[0]), [int(x) for x in s[1:]]


def _worker(rank, world, n_batches, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine()
    multigpu.run_sharded(eng, n_batches, rank, world, torch.device("cpu"))
    multigpu.reduce_accumulators([(eng.acc, "sum")], rank, world)
    multigpu.gather_variable(eng, rank, world, torch.device("cpu"))
    out[rank] = (dict(eng.log), eng.acc.clone(), dict(eng.var), dict(eng.got))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_batches,world,port", [(9, 2, 29531), (7, 3, 29541), (2, 4, 29551), (8, 4, 29561)])
def test_ring_hand_off_matches_single_process(n_batches, world, port):
    """Odd batch counts, a world that does not divide them, and more ranks than batches (idle ranks still take part in the
    hand-off ring, the reduce and the gather)."""
    single = FakeEngine()
    multigpu.run_sharded(single, n_batches, 0, 1, torch.device("cpu"))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, n_batches, port, out), nprocs=world, join=True)
    merged = {}
    for r in range(world):
        merged.update(out[r][0])
    assert merged == single.log
    assert torch.equal(out[0][1], single.acc)
    # rank 0 received exactly the variable-size state of the other ranks, in rank order
    # (a rank that holds nothing sends nothing)
    want = {w: [out[r][2][w] for r in range(1, world) if out[r][2][w]] for w in (0, 1)}
    assert {w: v for w, v in out[0][3].items() if v} == {w: v for w, v in want.items() if v}
    assert sorted(sum((out[r][2][1] for r in range(world)), [])) == sorted(single.var[1])
