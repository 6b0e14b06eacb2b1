"""world_size-2 gloo tests (CPU) of the repository's host-side multi-rank logic: the SyncBatchNorm statistics
exchange of the fused path (fused.all_reduce_stats: one [2C+1] fp64 buffer, row count reduced with the sums) and
bench.py's scene sharding / max-over-ranks timing helpers.  The kernels' side of SyncBatchNorm (device-side count,
rank-local dgamma/dbeta) is covered on the GPU by tests/test_gpu_syncbn.py."""
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
    # the layout pn2_bn_reduce_stats produces: (sum[C], sum of squares[C], rows of this rank)
    sums = torch.cat([mine.sum(0), (mine * mine).sum(0), torch.tensor([float(rows[rank])], dtype=torch.float64)])
    tot = fused.all_reduce_stats(sums, dist.group.WORLD)
    assert tot.data_ptr() == sums.data_ptr()  # in place: no host round trip, no re-allocation
    count = tot[24]
    mean = tot[:12] / count
    var = tot[12:24] / count - mean * mean
    ok = float(count) == 700.0 and torch.allclose(mean, full.mean(0)) and torch.allclose(var, full.var(0, unbiased=False))
    return bool(ok)


def test_syncbn_statistics_exchange_gloo(built_lib):
    out = _spawn(_syncbn_stats)
    assert out == {0: True, 1: True}


def _bench_sharding(rank, world):
    """bench.py: every rank draws its own scenes (scene_seed) and the step time is the max over ranks."""
    import bench
    ms = bench.max_over_ranks(10.0 + rank, torch.device("cpu"), world)
    return (ms, bench.scene_seed(rank))


def test_bench_sharding_and_max_reduce_gloo(built_lib):
    out = _spawn(_bench_sharding)
    assert out[0][0] == out[1][0] == 11.0
    assert out[0][1] != out[1][1]


def test_sync_group_detection(built_lib):
    """fused._sync_group: plain BatchNorm, eval mode or no process group -> no exchange."""
    import fused
    bn = torch.nn.SyncBatchNorm(4)
    assert fused._sync_group(bn) is None            # no process group initialised
    assert fused._sync_group(torch.nn.BatchNorm2d(4)) is None
    bn.eval()
    assert fused._sync_group(bn) is None
