"""CPU, world_size 2, gloo: the data-parallel training plumbing of BASELINE config 5 (bucketed gradient all-reduce,
identical SGD updates on every rank) on a toy torch model — the KPConv kernels themselves need a GPU (tests -m gpu)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apr_b200.train import GradBucketReducer, NPRHead, chamfer


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(),
                               torch.nn.Linear(64, 3), torch.nn.Linear(3, 3))     # last layer unused by the loss below


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(16, 8, generator=g), torch.randn(16, 3, generator=g)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _toy()
    red = GradBucketReducer(net.parameters(), bucket_mb=0.01)        # tiny buckets: several all-reduces per step
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-6)
    x, y = _data(rank)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = ((net[:5](x) - y) ** 2).mean()                        # net[5] gets no gradient: reduced as zeros
        loss.backward()
        red.finish()
        opt.step()
    q.put((rank, len(red.buckets), [p.detach().numpy().copy() for p in net.parameters()]))   # numpy: no fd passing
    dist.destroy_process_group()


def _worker_accum(rank, world, port, q):
    """iter_size = 2 (lib/trainer.py:316-322): two micro-batches per step, the first inside no_sync()."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _toy()
    red = GradBucketReducer(net.parameters(), bucket_mb=0.01)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    micro = [_data(10 * rank), _data(10 * rank + 1)]
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        with red.no_sync():
            (((net[:5](micro[0][0]) - micro[0][1]) ** 2).mean() / 2).backward()
        (((net[:5](micro[1][0]) - micro[1][1]) ** 2).mean() / 2).backward()
        red.finish()
        opt.step()
    q.put((rank, [p.detach().numpy().copy() for p in net.parameters()]))
    dist.destroy_process_group()


def test_two_rank_gradient_accumulation_no_sync():
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_accum, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    net = _toy()                                                       # single process: mean over the four micro-batches
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        loss = sum(((net[:5](x) - y) ** 2).mean() for x, y in (_data(0), _data(1), _data(10), _data(11))) / 4
        loss.backward()
        for p in net[5].parameters():
            p.grad = torch.zeros_like(p)
        opt.step()
    for a, b in zip(out[0][1], net.parameters()):
        assert np.allclose(a, b.detach().numpy(), atol=1e-6, rtol=1e-5)


def test_two_rank_bucketed_allreduce_matches_single_process_mean():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert out[0][1] > 1                                              # really bucketed
    import numpy as np
    for a, b in zip(out[0][2], out[1][2]):
        assert np.array_equal(a, b)                                    # identical replicas after the updates
    # single-process reference: gradient = mean of the two ranks' gradients
    net = _toy()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-6)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = sum(((net[:5](x) - y) ** 2).mean() for x, y in (_data(0), _data(1))) / 2
        loss.backward()
        for p in net[5].parameters():
            p.grad = torch.zeros_like(p)
        opt.step()
    for a, b in zip(out[0][2], net.parameters()):
        assert np.allclose(a, b.detach().numpy(), atol=1e-6, rtol=1e-5)


def test_npr_head_and_chamfer_semantics():
    head = NPRHead(in_channel=32, out_points=4)
    assert [m[0].out_features for m in head.list_modules] == [512, 256, 12]          # GenerativeMLP_98, mlp.py:156-157
    assert head(torch.randn(10, 32)).shape == (10, 12)
    a = torch.tensor([[0., 0, 0], [1, 0, 0]]); b = torch.tensor([[0., 0, 1], [1, 0, 0], [5, 0, 0]])
    # forward: (1 + 0)/2 ; backward: (1 + 0 + 16)/3
    assert abs(chamfer(a, b).item() - (0.5 + 17.0 / 3.0)) < 1e-6
