"""world_size-2 gloo test (CPU) of the N>1 path's host logic: sharding covers every image exactly once, the timing
reduction is a max over ranks, counts gather to whole-job totals — the same helpers bench.py uses under torchrun."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from focoos_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    b, e = D.shard_range(n_items)
    D.synchronize()
    slowest = D.max_over_ranks(10.0 + rank * 5.0)
    counts = D.gather_counts(e - b)
    q.put((rank, b, e, slowest, counts))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [64, 7, 1])
def test_two_rank_sharding_and_timing(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    covered = []
    for rank, b, e, slowest, counts in res:
        covered += list(range(b, e))
        assert slowest == 15.0, "timing must be the max over ranks"
        assert sum(counts) == n_items
    assert covered == list(range(n_items)), "every unit exactly once, no overlap"


def test_single_process_degenerates():
    assert D.get_world_size() == 1 and D.get_rank() == 0
    assert D.shard_range(10) == (0, 10)
    assert D.max_over_ranks(3.5) == 3.5 and D.gather_counts(4) == [4]
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
