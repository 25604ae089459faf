"""world_size-2..4 gloo test of the sharded run's dealing loop (fastquick_b200/multigpu.py) with a stand-in engine whose
hand-off and merge do over gloo what the C library does over NVLink / NCCL (fqb_collect_pairs_sharded: receive from the owner
of batch b-1 unless b == 0, send to the owner of b+1 unless b is the last; fqb_comm_merge_stats: reduce onto rank 0): the
state every batch starts from, and the merged totals, must be exactly those of a single process."""
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

    def __init__(self, rank=0, world=1):
        self.rank, self.world = rank, world
        self.calls, self.ii, self.log, self.acc = 0, [0] * 7, {}, torch.zeros(16, dtype=torch.int64)
        self.submitted, self.in_flight, self.first_pairs = [], [], {}

    def submit(self, b):
        assert len(self.in_flight) < 2, "at most two batches in flight"
        self.in_flight.append(b); self.submitted.append(b)

    def collect(self, b, first_pair, is_last):
        assert self.in_flight.pop(0) == b, "batches are collected in submission order"
        self.first_pairs[b] = first_pair
        if self.world > 1 and b > 0:
            buf = torch.zeros(8, dtype=torch.int64)
            dist.recv(buf, src=(b - 1) % self.world)
            self.calls, self.ii = int(buf[0]), [int(x) for x in buf[1:]]
        self.log[b] = (self.calls, tuple(self.ii))       # state this batch started from
        self.calls += 1000 + 17 * b
        if b % 3 != 2:                                    # inference "succeeds"
            self.ii = [b * 11 + k for k in range(7)]
        if self.world > 1 and not is_last:
            dist.send(torch.tensor([self.calls] + self.ii, dtype=torch.int64), dst=(b + 1) % self.world)
        self.acc[b % 16] += b + 1

    def merge(self):
        if self.world > 1:
            dist.reduce(self.acc, dst=0)


def _worker(rank, world, n_batches, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine(rank, world)
    mine = multigpu.run_sharded(eng, n_batches, rank, world, 1000)
    eng.merge()
    out[rank] = (dict(eng.log), eng.acc.clone(), list(mine), dict(eng.first_pairs))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_batches,world,port", [(9, 2, 29531), (7, 3, 29541), (2, 4, 29551), (8, 4, 29561)])
def test_ring_hand_off_matches_single_process(n_batches, world, port):
    """Odd batch counts, a world that does not divide them, and more ranks than batches (idle ranks only take part in the merge)."""
    single = FakeEngine()
    multigpu.run_sharded(single, n_batches, 0, 1, 1000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, n_batches, port, out), nprocs=world, join=True)
    merged, firsts = {}, {}
    for r in range(world):
        merged.update(out[r][0]); firsts.update(out[r][3])
        assert out[r][2] == [b for b in range(n_batches) if b % world == r]
    assert merged == single.log
    assert firsts == {b: b * 1000 for b in range(n_batches)}
    assert torch.equal(out[0][1], single.acc)
