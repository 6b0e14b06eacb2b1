import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'omni-pq_b200'))
import bench, _pn2
from backbone import Pointnet2Backbone
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Pointnet2Backbone(input_feature_dim=3).to(dev).train()
host = bench.make_scenes(4, 40000, 1234).pin_memory(); res = host.to(dev)
def step(it, sync=False):
    net.zero_grad(set_to_none=True)
    ep = net(res[it % 4][None]); ep["fp2_features"].sum().backward()
    if sync: torch.cuda.synchronize()
for it in range(3): step(it)
torch.cuda.synchronize()
for rep in range(4):
    for sync in (False, True):
        t0 = time.perf_counter(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for it in range(10): step(it, sync)
        t_cpu = time.perf_counter() - t0
        e.record(); torch.cuda.synchronize()
        print(f"rep {rep} sync={sync}: gpu {s.elapsed_time(e)/10:.2f} ms/step, cpu enqueue {t_cpu*100:.2f} ms/step, mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB reserved {torch.cuda.memory_reserved()/2**30:.2f}", flush=True)
