"""Worker of tests/test_gpu_syncbn.py (launched twice by torchrun, both ranks on cuda:0, gloo backend -- NCCL refuses
two ranks on one device, gloo all-reduces CUDA tensors through the host, which is all this semantic check needs).

Ours: fused PointnetSAModuleVotes / PointnetFPModule whose BatchNorm children were converted by
nn.SyncBatchNorm.convert_sync_batchnorm (what models/pq_transformer.py:194 does), wrapped in DDP
(broadcast_buffers=False, train.py:382).  Reference: the oracle's module glue over the reference's own CUDA kernels
(oracle/_ref) with torch.nn.SyncBatchNorm + DDP.  Each rank sees different clouds; outputs, rank-averaged parameter
gradients, input gradients and running statistics must agree."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "omni-pq_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = os.environ.get("PN2_SYNCBN_BACKEND", "gloo")
    dev = int(os.environ.get("LOCAL_RANK", "0")) if backend == "nccl" else 0  # nccl: one GPU per rank; gloo: both on cuda:0
    torch.cuda.set_device(dev)
    import datetime
    dist.init_process_group(backend, timeout=datetime.timedelta(seconds=120))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import pointnet2_modules as M
    import _pn2
    from oracle import build_ref, pn2_oracle as O
    ext = build_ref.load()
    assert ext is not None, "oracle/_ref/pn2_ref_ext.so missing"
    O.ext = ext  # the oracle's glue on the reference's CUDA kernels

    class Pair(torch.nn.Module):  # SA followed by FP back onto the input points: both fused paths, chained
        def __init__(self, sa_cls, fp_cls):
            super().__init__()
            self.sa = sa_cls(npoint=96, radius=0.3, nsample=16, mlp=[8, 24, 32], use_xyz=True, normalize_xyz=True)
            self.fp = fp_cls(mlp=[32 + 8, 24, 16])

        def forward(self, xyz, feats):
            new_xyz, f, _ = self.sa(xyz, feats)
            return self.fp(xyz, new_xyz, feats, f), f

    torch.manual_seed(3)
    ours = Pair(M.PointnetSAModuleVotes, M.PointnetFPModule)
    theirs = Pair(O.OracleSAModuleVotes, O.OracleFPModule)
    theirs.load_state_dict(ours.state_dict(), strict=True)
    ours = torch.nn.SyncBatchNorm.convert_sync_batchnorm(ours).cuda().train()
    theirs = torch.nn.SyncBatchNorm.convert_sync_batchnorm(theirs).cuda().train()
    assert isinstance(ours.sa.mlp_module.layer0.bn.bn, torch.nn.SyncBatchNorm)
    d_ours = DDP(ours, device_ids=[dev], broadcast_buffers=False)
    d_theirs = DDP(theirs, device_ids=[dev], broadcast_buffers=False)

    n = [500, 500][rank]
    b = [2, 3][rank]  # ranks hold different numbers of rows: the global count matters
    xyz, feats = O.uniform_cloud(b, n, 8, seed=40 + rank)
    xyz = xyz.cuda()
    f1, f2 = feats.cuda().requires_grad_(True), feats.cuda().requires_grad_(True)
    before = _pn2.launch_count
    out, mid = d_ours(xyz, f1)
    assert _pn2.launch_count - before >= 10, "fused path did not run"
    out_t, mid_t = d_theirs(xyz, f2)
    errs = {"sa_out": rel(mid, mid_t), "fp_out": rel(out, out_t)}
    cot = torch.randn(out_t.shape, generator=torch.Generator().manual_seed(7 + rank)).cuda()
    (out * cot).sum().backward()
    (out_t * cot).sum().backward()
    errs["d_feats"] = rel(f1.grad, f2.grad)
    worst_name, worst = None, 0.0
    for (n1, p1), (n2, p2) in zip(ours.named_parameters(), theirs.named_parameters()):
        assert n1 == n2
        r = rel(p1.grad, p2.grad)
        if r > worst:
            worst_name, worst = n1, r
    errs["param_grad"] = worst
    bw = 0.0
    for (n1, b1), (n2, b2) in zip(ours.named_buffers(), theirs.named_buffers()):
        if b1.dtype.is_floating_point:
            bw = max(bw, rel(b1, b2))
        else:
            assert torch.equal(b1, b2), n1
    errs["buffers"] = bw
    ok = errs["sa_out"] <= 1e-5 and errs["fp_out"] <= 1e-5 and errs["d_feats"] <= 1e-4 and worst <= 1e-4 and bw <= 1e-5
    import fused
    why = [ex.why for ex in fused._peer_exchanges.values() if not ex.ok]
    print(f"rank {rank}: {'SYNCBN_OK' if ok else 'SYNCBN_FAIL'} {errs} worst_param={worst_name} "
          f"peer_exchanges={fused.peer_exchange_calls} {why}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
