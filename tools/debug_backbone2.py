import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'omni-pq_b200'))
from oracle import pn2_oracle as O
import _pn2 as K
from backbone import Pointnet2Backbone
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
torch.manual_seed(0)
ours = Pointnet2Backbone(input_feature_dim=3)
oracle = O.OracleBackbone(input_feature_dim=3)
oracle.load_state_dict(ours.state_dict())
ours.cuda().train(); oracle.train()
cloud = O.scannet_like_cloud(npts, seed=1234)[None]
def run(net, pc):
    t = {}
    xyz = pc[..., :3].contiguous(); f = pc[..., 3:].transpose(1, 2).contiguous()
    x1, f1, _ = net.sa1(xyz, f); x2, f2, _ = net.sa2(x1, f1); x3, f3, _ = net.sa3(x2, f2); x4, f4, _ = net.sa4(x3, f3)
    g1 = net.fp1(x3, x4, f3, f4); g2 = net.fp2(x2, x3, f2, g1)
    for k, v in dict(f1=f1, f2=f2, f3=f3, f4=f4, g1=g1, g2=g2).items():
        v.retain_grad(); t[k] = v
    return t
a = run(ours, cloud.cuda()); b = run(oracle, cloud)
cot = torch.randn(b["g2"].shape, generator=torch.Generator().manual_seed(1))
(a["g2"] * cot.cuda()).sum().backward(); (b["g2"] * cot).sum().backward()
for k in ["g2", "g1", "f4", "f3", "f2", "f1"]:
    e = a[k].grad.cpu() - b[k].grad
    print(k, "fwd", f"{rel(a[k], b[k]):.2e}", "grad rel", f"{rel(a[k].grad, b[k].grad):.2e}", " per-channel mean err max", float(e.mean(dim=2).abs().max()), " ref max", float(b[k].grad.abs().max()),
          " #elements off by >1e-3*max:", int((e.abs() > 1e-3 * b[k].grad.abs().max()).sum()), "of", e.numel())
