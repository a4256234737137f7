"""N>1 host logic under gloo (world_size 2, CPU): stream sharding, the barrier
around the timed region and the max-over-ranks reduction used by bench.py."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from joshupscale_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_streams, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.streams_for_rank(total_streams, world, rank)
    dist.barrier()
    elapsed = 0.010 * (rank + 1)  # rank 1 is the slow one
    slowest = sharding.max_over_ranks(elapsed, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((gathered, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_max_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered == [[0, 2, 4, 6], [1, 3, 5]]
    assert sorted(sum(gathered, [])) == list(range(7))  # every stream exactly once
    assert abs(slowest - 0.020) < 1e-9
    assert sharding.aggregate_fps(4, 2, 100, slowest) == pytest.approx(4 * 2 * 100 / 0.020)


def test_sharding_edge_cases():
    assert sharding.streams_for_rank(0, 4, 1) == []
    assert sharding.streams_for_rank(3, 8, 5) == []
    assert sharding.streams_for_rank(64, 8, 7) == list(range(7, 64, 8))
    assert sharding.max_over_ranks(1.5) == 1.5
    with pytest.raises(ValueError):
        sharding.streams_for_rank(4, 2, 2)
