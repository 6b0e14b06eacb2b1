"""Worker of tests/test_gpu_multi.py (torchrun, one GPU per rank, NCCL): graphed.GraphedTrainStep with a process group
(forward + backward + in-graph NCCL all-reduce of the gradient arena, ReduceOp.AVG) must leave in every param.grad what
DistributedDataParallel leaves there for the same model and the same per-rank inputs (train.py:382)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "omni-pq_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    from backbone import Pointnet2Backbone
    from graphed import GraphedTrainStep
    from tools.synth_clouds import scannet_like_cloud

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = Pointnet2Backbone(input_feature_dim=3)

        def forward(self, cloud):
            return self.backbone(cloud)["fp2_features"]

    torch.manual_seed(0)
    net = Net().cuda().train()
    state0 = {k: v.clone() for k, v in net.state_dict().items()}
    cloud = scannet_like_cloud(12000, seed=500 + rank)[None].cuda()   # every rank its own scene
    cot = torch.randn(1, 288, 1024, generator=torch.Generator().manual_seed(9 + rank)).cuda()
    # reference: eager DDP
    ddp = DDP(net, device_ids=[local], broadcast_buffers=False)
    (ddp(cloud) * cot).sum().backward()
    want = [p.grad.detach().clone() for p in net.parameters()]
    net.load_state_dict(state0)
    net.zero_grad(set_to_none=True)
    del ddp
    step = GraphedTrainStep(net, lambda out: (out * cot).sum(), (cloud,), process_group=dist.group.WORLD)
    for _ in range(2):
        net.load_state_dict(state0)
        step(cloud)
    torch.cuda.synchronize()
    worst, name = 0.0, None
    for (n, p), w in zip(net.named_parameters(), want):
        d = float((p.grad - w).abs().max() / w.abs().max().clamp_min(1e-30))
        if d > worst:
            worst, name = d, n
    ok = worst <= 5e-5
    print(f"rank {rank}: {'GRAPH_DDP_OK' if ok else 'GRAPH_DDP_FAIL'} worst {worst:.3e} at {name}; launches/step {step.launches_per_step}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
