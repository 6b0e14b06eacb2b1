"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: the SyncBatchNorm statistics exchange
used by the fused path, scene sharding of the benchmark harness, and DDP gradient averaging over the
reference-named parameters."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return dict(ret)


def _syncbn_stats(rank, world):
    import fused
    g = torch.Generator().manual_seed(100)
    full = torch.randn(700, 12, generator=g).double() * 3 + 1.5
    rows = [300, 400]
    start = sum(rows[:rank])
    mine = full[start:start + rows[rank]]
    sums = torch.stack([mine.sum(0), (mine * mine).sum(0)])
    tot, count = fused.all_reduce_stats(sums, rows[rank], dist.group.WORLD)
    mean = tot[0] / count
    var = tot[1] / count - mean * mean
    ok = count == 700.0 and torch.allclose(mean, full.mean(0)) and torch.allclose(var, full.var(0, unbiased=False))
    return bool(ok)


def test_syncbn_statistics_exchange_gloo(built_lib):
    out = _spawn(_syncbn_stats)
    assert out == {0: True, 1: True}


def _ddp_grad_average(rank, world):
    """DDP over the reference-named SharedMLP parameters: gradients are averaged across ranks (what
    train.py:382 relies on).  CPU tensors go through the module's op-level fallback only at the nn level."""
    import pytorch_utils as P
    torch.manual_seed(0)
    mlp = P.SharedMLP([6, 8, 4], bn=True)
    ddp = torch.nn.parallel.DistributedDataParallel(mlp, broadcast_buffers=False)
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(2, 6, 5, 3, generator=g)
    ddp(x).sum().backward()
    w = mlp.layer0.conv.weight.grad.clone()
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    return bool(torch.allclose(gathered[0], gathered[1]))


def test_ddp_gradient_allreduce_gloo(built_lib):
    out = _spawn(_ddp_grad_average)
    assert out == {0: True, 1: True}


def _bench_sharding(rank, world):
    """bench.py gives every rank its own scenes (seed offset by rank) and reports max-over-ranks time."""
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    seed0 = 1234 + 100 * rank
    return (float(ms.item()), seed0)


def test_bench_sharding_and_max_reduce_gloo(built_lib):
    out = _spawn(_bench_sharding)
    assert out[0][0] == out[1][0] == 11.0
    assert out[0][1] != out[1][1]
