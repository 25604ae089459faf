"""Launcher-side plumbing of a sharded run: one process per GPU, `torch.distributed` only to bootstrap.

Everything that touches data is in the C library: batches of 262,144 pairs go round-robin to ranks (batch b -> rank
b % world) with the index replicated; `fqb_collect_pairs_sharded` receives the drand48 position + `last_ii` the owner of
batch b-1 left (56 bytes, written into this rank's mailbox over NVLink peer memory) and hands its own on;
`fqb_comm_merge_stats` reduces the accumulators onto rank 0 with NCCL and moves the pile-up entries / duplicate keys there.
What is left for Python is the rendezvous (`bootstrap`: the NCCL id of rank 0 and every rank's mailbox handle travel over
the process group the launcher already has) and the loop that deals the batches (`run_sharded`), which is written against a
small engine interface so that it can be exercised on CPU with gloo (tests/test_multigpu_protocol.py).
"""
import ctypes as C

import torch
import torch.distributed as dist


def bootstrap(lib, h, rank, world):
    """fqb_comm_init for handle `h`: exchange rank 0's NCCL unique id and all mailbox handles over torch.distributed."""
    if world == 1:
        assert lib.fqb_comm_init(h, 0, 1, None, None) == 0, lib.fqb_last_error()
        return
    box = (C.c_uint8 * 64)()
    assert lib.fqb_comm_ring_handle(h, box) == 0, lib.fqb_last_error()
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        assert lib.fqb_comm_unique_id(uid) == 0, lib.fqb_last_error()
    ids = [bytes(uid)]
    dist.broadcast_object_list(ids, src=0)
    boxes = [None] * world
    dist.all_gather_object(boxes, bytes(box))
    all_boxes = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(boxes))
    uid = (C.c_uint8 * 128).from_buffer_copy(ids[0])
    assert lib.fqb_comm_init(h, rank, world, uid, all_boxes) == 0, lib.fqb_last_error()


def run_sharded(engine, n_batches, rank, world, batch_pairs):
    """Deal `n_batches` batches of one file over the ranks and drive this rank's share through the pipelined calls.

    engine.submit(b)                                  -- upload + align stage of global batch b (asynchronous)
    engine.collect(b, first_pair, is_last)            -- pairing .. statistics of batch b, with the hand-off on either side
    Batch b+world is submitted before batch b is collected, so its align stage overlaps the later stages of b.
    """
    mine = [b for b in range(n_batches) if b % world == rank]
    if mine:
        engine.submit(mine[0])
    for j, b in enumerate(mine):
        if j + 1 < len(mine):
            engine.submit(mine[j + 1])
        engine.collect(b, b * batch_pairs, b == n_batches - 1)
    return mine
