"""CPU, world_size 2, gloo: the N > 1 host logic (pair sharding, MAX-over-ranks timing, SUM of units)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apr_b200.shard import aggregate_throughput, shard_indices


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(11, rank, world)
    secs = 2.0 if rank == 0 else 3.0                      # rank 1 is the slow one
    thr, n, t = aggregate_throughput(2 * len(mine), secs)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, thr, n, t, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    all_items = sorted(out[0][1] + out[1][1])
    assert all_items == list(range(11))                   # every pair processed exactly once
    assert set(out[0][1]).isdisjoint(out[1][1])
    for _, _, thr, n, t, gathered in out:
        assert n == 22.0 and t == 3.0 and abs(thr - 22.0 / 3.0) < 1e-12      # SUM of clouds / MAX of seconds
        assert gathered == [out[0][1], out[1][1]]


def test_single_process_fallback():
    thr, n, t = aggregate_throughput(10, 2.0)
    assert thr == 5.0 and n == 10.0 and t == 2.0
    assert shard_indices(5, 1, 2) == [1, 3]
