"""Multi-GPU orchestration of the align stage: one process per GPU, `torch.distributed` for the plumbing.

Batches of 262,144 pairs go round-robin to ranks (batch b -> rank b % world); the index is replicated.  The hot path has
no data-path collective.  What does cross ranks:

  * between `fqb_stage_align` and `fqb_stage_pair` of batch b, the owner of batch b-1 sends the position of the global
    drand48 stream (draws consumed so far) and its insert-size estimate (`last_ii`): 8 + 48 bytes, point to point;
  * at the end, the integer accumulators are reduced to rank 0 (sum; first-touch order: min), and the variable-size
    state -- marker pile-up entries (tagged with their global pair index) and the distinct PCR-duplicate keys --
    is gathered there; rank 0 then splices the ranks' InsertSizeTable batches into file order and writes the files.

The engine is any object with the small interface used below, so the protocol is testable on CPU with gloo
(tests/test_multigpu_protocol.py) and runs on NCCL in bench.py.
"""
import torch
import torch.distributed as dist

STATE_WORDS = 8    # rng_calls + isize_info_t (avg, std, ap_prior as f64 bit patterns; low, high, high_bayesian, pad as u32 pairs)


def run_sharded(engine, n_batches, rank, world, device):
    """Drive `engine` over this rank's batches with the cross-batch state handed along the ring.

    engine.align(b), engine.pair(b), engine.finish(b)     -- per-batch stages (finish = SW/refine + stats)
    engine.get_state() -> list[int] (STATE_WORDS int64)   -- after pair(b)
    engine.set_state(list[int])                           -- before pair(b)
    """
    mine = [b for b in range(n_batches) if b % world == rank]
    for b in mine:
        engine.align(b)                                     # heavy, no dependency on other batches
        if b > 0:
            src = (b - 1) % world
            if src != rank:
                buf = torch.zeros(STATE_WORDS, dtype=torch.int64, device=device)
                dist.recv(buf, src=src)
                engine.set_state(buf.tolist())
        engine.pair(b)
        if b + 1 < n_batches:
            dst = (b + 1) % world
            if dst != rank:
                dist.send(torch.tensor(engine.get_state(), dtype=torch.int64, device=device), dst=dst)
        engine.finish(b)
    return mine


def reduce_accumulators(groups, rank, world):
    """groups: list of (tensor, op) living on this rank's device; reduced in place onto rank 0."""
    if world == 1:
        return
    for t, op in groups:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN)


VAR_ITEM_BYTES = {0: 20, 1: 8}     # pile-up entries, duplicate keys (fqb_stats_var_*)


def gather_variable(engine, rank, world, device):
    """Gather the variable-size statistics state onto rank 0 (after reduce_accumulators + import of the sums).

    engine.var_export(which) -> 1-D uint8 torch tensor (on `device` or on the host);  engine.var_import(which, uint8
    tensor on `device`) on rank 0.
    """
    if world == 1:
        return
    import os
    import time
    verbose = bool(os.environ.get("FQB_BENCH_VERBOSE")) and rank == 0 and device.type == "cuda"
    for which in sorted(VAR_ITEM_BYTES):
        t0 = time.time()
        mine = engine.var_export(which)
        if verbose:
            torch.cuda.synchronize(); t1 = time.time()
        n = torch.tensor([mine.numel()], dtype=torch.int64, device=device)
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        cap = max(max(sizes), 1)
        buf = torch.zeros(cap, dtype=torch.uint8, device=device)
        buf[: mine.numel()] = mine.to(device)
        out = [torch.zeros(cap, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0)
        if verbose:
            torch.cuda.synchronize(); t2 = time.time()
        if rank == 0:
            for r in range(1, world):
                if sizes[r]:
                    engine.var_import(which, out[r][: sizes[r]])
        if verbose:
            torch.cuda.synchronize(); t3 = time.time()
            print("  var group %d: export %.1f ms, gather %.1f ms (%d bytes max), import %.1f ms" % (
                which, (t1 - t0) * 1e3, (t2 - t1) * 1e3, cap, (t3 - t2) * 1e3), flush=True)
