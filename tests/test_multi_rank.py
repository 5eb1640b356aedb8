"""The N>1 path of bench.py on CPU: two ranks over gloo.  The data path has no collective (jobs are independent); what
is distributed is the job split and the max-over-ranks timing."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from fastsmc_b200 import asmc
    jobs = asmc.pyASMC.jobsOfRank(16, world, rank)
    t_ms = bench.max_over_ranks(100.0 * (rank + 1), world, device="cpu")
    total = bench.sum_over_ranks(10 + rank, world, device="cpu")
    # the dynamic job queue of bench.jobs_run: a counter in the process group's store hands every job out exactly once
    order = asmc.pyASMC.jobOrder(16)
    store = dist.distributed_c10d._get_default_store()
    taken = []
    while True:
        i = store.add("fsmc_jobs_next", 1) - 1
        if i >= len(order):
            break
        taken.append(order[i])
    gathered = [None] * world
    dist.all_gather_object(gathered, (jobs, taken))
    if rank == 0:
        out.put(([g[0] for g in gathered], t_ms, total, [g[1] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_split_jobs_and_take_max_time():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, t_ms, total, taken = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(1, 17))  # every job exactly once
    assert not set(gathered[0]) & set(gathered[1])
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert t_ms == 200.0  # the slower rank defines the step time
    assert total == 21.0
    assert sorted(taken[0] + taken[1]) == list(range(1, 17))  # shared counter: every job once, whatever the interleaving


def test_job_order_puts_the_remainder_job_first():
    sys.path.insert(0, ROOT)
    from fastsmc_b200 import asmc
    for J in (1, 4, 9, 16, 64):
        order = asmc.pyASMC.jobOrder(J)
        assert sorted(order) == list(range(1, J + 1)) and order[0] == J


def test_batch_shares_partition_the_job():
    """bench.py under torchrun: rank r decodes the r-th contiguous share of the job's reference batches."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    a, b = bench.all_pairs_in_reference_order(40)  # 3 160 pairs -> 99 batches, the last one partial
    whole = bench.tiles_for(a, b, 500)
    for world in (1, 2, 3, 8):
        parts = [bench.tiles_for(a, b, 500, world, r) for r in range(world)]
        assert sum(len(t["tilePairs"]) for t in parts) == len(whole["tilePairs"])
        assert np.array_equal(np.concatenate([t["hapA"] for t in parts]), whole["hapA"])
        assert np.array_equal(np.concatenate([t["hapB"] for t in parts]), whole["hapB"])
        assert np.array_equal(np.concatenate([t["tilePairs"] for t in parts]), whole["tilePairs"])
        sizes = [len(t["tilePairs"]) for t in parts]
        assert max(sizes) - min(sizes) <= 1
