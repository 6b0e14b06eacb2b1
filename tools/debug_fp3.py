import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'omni-pq_b200'))
from oracle import pn2_oracle as O
import _pn2 as K, pointnet2_modules as M
import test_gpu_fused as T
def run(n, m, c1, c2, mlp):
    ours, oracle = T._pair(lambda: M.PointnetFPModule(mlp=list(mlp)), lambda: O.OracleFPModule(mlp=list(mlp)), seed=7)
    ours.train(); oracle.train()
    unknown, uf = O.uniform_cloud(1, n, c1, seed=21)
    known, kf = O.uniform_cloud(1, m, c2, seed=22)
    uf_d, kf_d = uf.cuda().requires_grad_(True), kf.cuda().requires_grad_(True)
    uf_c, kf_c = uf.clone().requires_grad_(True), kf.clone().requires_grad_(True)
    out = ours(unknown.cuda(), known.cuda(), uf_d, kf_d); out_o = oracle(unknown, known, uf_c, kf_c)
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(1))
    (out * cot.cuda()).sum().backward(); (out_o * cot).sum().backward()
    print(f"n={n} m={m} c1={c1} c2={c2} mlp={mlp}: out {T.rel(out, out_o):.2e} d_unknow {T.rel(uf_d.grad, uf_c.grad):.2e} d_known {T.rel(kf_d.grad, kf_c.grad):.2e}",
          " params", " ".join(f"{T.rel(p1.grad, p2.grad):.1e}" for p1, p2 in zip(ours.parameters(), oracle.parameters())))
    e = (kf_d.grad.cpu() - kf_c.grad)
    print("   d_known err: per-channel mean of err / max", float(e.mean(dim=2).abs().max()), float(e.abs().max()), " ref max", float(kf_c.grad.abs().max()))
run(900, 150, 12, 64, [76, 48, 20])
run(1024, 512, 512, 512, [1024, 512, 288])
run(1024, 512, 64, 64, [128, 64, 32])
run(1024, 500, 64, 512, [576, 64, 32])
run(512, 256, 512, 512, [1024, 512, 512])
