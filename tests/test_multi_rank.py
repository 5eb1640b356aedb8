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
    gathered = [None] * world
    dist.all_gather_object(gathered, jobs)
    if rank == 0:
        out.put((gathered, t_ms))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_split_jobs_and_take_max_time():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, t_ms = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(1, 17))  # every job exactly once
    assert not set(gathered[0]) & set(gathered[1])
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert t_ms == 200.0  # the slower rank defines the step time


def test_job_order_puts_the_remainder_job_first():
    sys.path.insert(0, ROOT)
    from fastsmc_b200 import asmc
    for J in (1, 4, 9, 16, 64):
        order = asmc.pyASMC.jobOrder(J)
        assert sorted(order) == list(range(1, J + 1)) and order[0] == J
