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
ep = ours(cloud.cuda()); ep_o = oracle(cloud)
for k in ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"]:
    print(k, rel(ep[k], ep_o[k]))
cot = torch.randn(ep_o["fp2_features"].shape, generator=torch.Generator().manual_seed(1))
(ep["fp2_features"] * cot.cuda()).sum().backward()
(ep_o["fp2_features"] * cot).sum().backward()
for (n, p1), (_, p2) in zip(ours.named_parameters(), oracle.named_parameters()):
    print(f"{n:45s} rel {rel(p1.grad, p2.grad):.3e}  |g|max {float(p2.grad.abs().max()):.3e}")
