import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'omni-pq_b200'))
from oracle import pn2_oracle as O
import _pn2 as K, fused, pointnet2_modules as M
import test_gpu_fused as T
torch.manual_seed(7)
ours = M.PointnetFPModule(mlp=[76, 48, 20]); T._randomise_bn(ours)
unknown, uf = O.uniform_cloud(2, 900, 12, seed=21)
known, kf = O.uniform_cloud(2, 150, 64, seed=22)
dist, idx = O.three_nn(unknown, known)
X = torch.cat([O.three_interpolate(kf, idx, O.fp_weights(dist)), uf], 1)
x = X.permute(0, 2, 1).reshape(-1, 76).contiguous()
ours.cuda().train()
layers = fused._mlp_layers(ours.mlp)
rows0 = K.rows_plain(x.cuda(), 1800, 76, 76)
state = fused._run_mlp(layers, rows0, 1800, 0, 0, False)
xc = x
for i, L in enumerate(state):
    conv, bn = layers[i]
    W = conv.weight.detach().cpu().view(L.cout, L.cin)
    y = xc @ W.t()
    print(f"layer {i}: y rel", float((L.y.cpu() - y).abs().max() / y.abs().max()))
    mu = y.double().mean(0); var = y.double().var(0, unbiased=False)
    invstd = 1 / torch.sqrt(var + bn.eps)
    print("   mean err", float((L.mean.cpu().double() - mu).abs().max()), " invstd rel err", float(((L.invstd.cpu().double() - invstd) / invstd).abs().max()))
    sc = bn.weight.detach().cpu().double() * invstd; sh = bn.bias.detach().cpu().double() - mu * sc
    print("   scale rel err", float(((L.scale.cpu().double() - sc) / sc).abs().max()), " shift err", float((L.shift.cpu().double() - sh).abs().max()))
    xc = torch.relu(y.double() * sc + sh).float()

print("---- X built by the kernels")
known_pm = K.to_point_major(kf.cuda())
xg = K._f32("cuda", 1800, 76)
idx_g, w_g = K.fp_interpolate(unknown.cuda(), known.cuda(), known_pm, 64, xg, 76)
K.to_point_major(uf.cuda(), ld=12, out=xg, col0=64)
print("idx equal", torch.equal(idx_g.cpu(), idx), " w rel", float((w_g.cpu() - O.fp_weights(dist)).abs().max()))
print("X rel", float((xg.cpu() - x).abs().max() / x.abs().max()), " interp part", float((xg.cpu()[:, :64] - x[:, :64]).abs().max()),
      " skip part", float((xg.cpu()[:, 64:] - x[:, 64:]).abs().max()))
torch.manual_seed(7)
ours2 = M.PointnetFPModule(mlp=[76, 48, 20]); T._randomise_bn(ours2)
oracle = O.OracleFPModule(mlp=[76, 48, 20]); oracle.load_state_dict(ours2.state_dict()); oracle.train()
ours2.cuda().train()
out = ours2(unknown.cuda(), known.cuda(), uf.cuda(), kf.cuda())
out_o = oracle(unknown, known, uf, kf)
d = (out.cpu() - out_o).abs()
print("module out rel", float(d.max() / out_o.abs().max()), "per-channel max err", d.amax(dim=(0, 2)))
print("oracle per-channel max", out_o.abs().amax(dim=(0, 2)))
