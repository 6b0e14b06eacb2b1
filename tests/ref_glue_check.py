"""Pin the oracle's Python glue against the reference's OWN pointnet2/*.py (run as a script by
tests/test_ref_glue_cpu.py, in a fresh process because the reference's modules carry the same top-level names
as ours).

The reference's unchanged `pointnet2_modules.py` / `pointnet2_utils.py` / `pytorch_utils.py` are imported from
/root/reference (or the staged copy baseline/_ref) with the CPU oracle's `ext` installed as `pointnet2._ext`
(the nine C functions, pinned bit-exact against the reference's CUDA kernels by the golden vectors).  On identical
CPU inputs and identical parameters the reference modules and the oracle's restatements
(`OracleSAModuleVotes`, `OracleFPModule`, `query_and_group`, `OracleBackbone`) must then agree BIT FOR BIT in
forward outputs, indices and every gradient -- which makes the float-path oracle "the reference's glue + torch CPU
fp32" rather than a builder restatement (pointnet2_modules.py:210-272,371-416,
pointnet2_utils.py:294-376).
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import pn2_oracle as O  # noqa: E402
from tools import stage_reference  # noqa: E402

ref = stage_reference.root()
assert ref is not None, "reference glue not available"
pkg = types.ModuleType("pointnet2")
pkg.__path__ = []
pkg._ext = O.ext
sys.modules["pointnet2"] = pkg
sys.modules["pointnet2._ext"] = O.ext
sys.path.insert(0, os.path.join(ref, "pointnet2"))
sys.path.insert(1, os.path.join(ref, "models"))

import pointnet2_modules as RM  # noqa: E402  (the reference's file)
import pointnet2_utils as RU  # noqa: E402

assert os.path.abspath(RM.__file__).startswith(ref), RM.__file__
torch.set_num_threads(4)


def same(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape)
    assert torch.equal(a, b), (what, float((a.double() - b.double()).abs().max()))


def grads_same(m1, m2, what):
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2, (n1, n2)
        same(p1.grad, p2.grad, f"{what}: grad {n1}")
    for (n1, b1), (n2, b2) in zip(m1.named_buffers(), m2.named_buffers()):
        same(b1, b2, f"{what}: buffer {n1}")


checks = 0

# ---- QueryAndGroup (pointnet2_utils.py:294-376) -----------------------------------------------------------------
xyz, feats = O.uniform_cloud(2, 512, 5, seed=3)
inds = O.ext.furthest_point_sampling(xyz, 64)
new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
for normalize in (False, True):
    for use_xyz in (True, False):
        q = RU.QueryAndGroup(0.25, 16, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize)
        got, got_xyz = q(xyz, new_xyz, feats)
        want, want_xyz, _ = O.query_and_group(0.25, 16, xyz, new_xyz, feats, use_xyz=use_xyz, normalize_xyz=normalize)
        same(got, want, f"QueryAndGroup normalize={normalize} use_xyz={use_xyz}")
        same(got_xyz, want_xyz, "grouped_xyz")
        checks += 2

# ---- PointnetSAModuleVotes fwd + bwd incl. the xyz gradient (pointnet2_modules.py:164-272) ------------------------
for train in (True, False):
    kw = dict(npoint=128, radius=0.3, nsample=16, use_xyz=True, normalize_xyz=True)
    torch.manual_seed(2)
    r_sa = RM.PointnetSAModuleVotes(mlp=[20, 32, 24, 40], **kw)
    o_sa = O.OracleSAModuleVotes(mlp=[20, 32, 24, 40], **kw)
    o_sa.load_state_dict(r_sa.state_dict(), strict=True)
    r_sa.train(train)
    o_sa.train(train)
    xyz, feats = O.uniform_cloud(3, 700, 20, seed=11)
    xr, xo = xyz.clone().requires_grad_(True), xyz.clone().requires_grad_(True)
    fr, fo = feats.clone().requires_grad_(True), feats.clone().requires_grad_(True)
    nx_r, out_r, inds_r = r_sa(xr, fr)
    nx_o, out_o, inds_o = o_sa(xo, fo)
    same(inds_r, inds_o, "SA inds")
    same(nx_r, nx_o, "SA new_xyz")
    same(out_r, out_o, "SA features")
    g = torch.Generator().manual_seed(3)
    cot, cot_xyz = torch.randn(out_o.shape, generator=g), torch.randn(nx_o.shape, generator=g)
    ((out_r * cot).sum() + (nx_r * cot_xyz).sum()).backward()
    ((out_o * cot).sum() + (nx_o * cot_xyz).sum()).backward()
    same(fr.grad, fo.grad, "SA dfeatures")
    same(xr.grad, xo.grad, "SA dxyz")
    grads_same(r_sa, o_sa, f"SA train={train}")
    checks += 6

# given indices, no features (C0 = 0, the train.sh default)
kw = dict(npoint=64, radius=0.25, nsample=8, use_xyz=True, normalize_xyz=False)
torch.manual_seed(0)
r_sa = RM.PointnetSAModuleVotes(mlp=[0, 16, 16], **kw)
o_sa = O.OracleSAModuleVotes(mlp=[0, 16, 16], **kw)
o_sa.load_state_dict(r_sa.state_dict(), strict=True)
xyz, _ = O.uniform_cloud(2, 333, 3, seed=4)
given = torch.randint(0, 333, (2, 64), generator=torch.Generator().manual_seed(0), dtype=torch.int32)
a, b = r_sa(xyz, None, given), o_sa(xyz, None, given)
same(a[0], b[0], "SA(given inds) new_xyz")
same(a[1], b[1], "SA(given inds) features")
checks += 2

# ---- PointnetFPModule fwd + bwd (pointnet2_modules.py:356-416) ------------------------------------------------------
for train in (True, False):
    torch.manual_seed(7)
    r_fp = RM.PointnetFPModule(mlp=[64 + 12, 48, 20])
    o_fp = O.OracleFPModule(mlp=[64 + 12, 48, 20])
    o_fp.load_state_dict(r_fp.state_dict(), strict=True)
    r_fp.train(train)
    o_fp.train(train)
    unknown, uf = O.uniform_cloud(2, 900, 12, seed=21)
    known, kf = O.uniform_cloud(2, 150, 64, seed=22)
    ur, uo = uf.clone().requires_grad_(True), uf.clone().requires_grad_(True)
    kr, ko = kf.clone().requires_grad_(True), kf.clone().requires_grad_(True)
    out_r, out_o = r_fp(unknown, known, ur, kr), o_fp(unknown, known, uo, ko)
    same(out_r, out_o, "FP features")
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(1))
    (out_r * cot).sum().backward()
    (out_o * cot).sum().backward()
    same(ur.grad, uo.grad, "FP d unknow_feats")
    same(kr.grad, ko.grad, "FP d known_feats")
    grads_same(r_fp, o_fp, f"FP train={train}")
    checks += 4

# ---- the reference's Pointnet2Backbone (models/backbone_module.py:33-139) vs OracleBackbone -------------------------
from backbone_module import Pointnet2Backbone  # noqa: E402  (the reference's file)

torch.manual_seed(0)
r_bb = Pointnet2Backbone(input_feature_dim=3).train()
o_bb = O.OracleBackbone(input_feature_dim=3).train()
o_bb.load_state_dict(r_bb.state_dict(), strict=True)
cloud = O.scannet_like_cloud(6000, seed=1234)[None]
ep_r, ep_o = r_bb(cloud), o_bb(cloud)
for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "sa1_features", "sa2_features",
          "sa3_features", "sa4_features", "fp2_features", "fp2_xyz", "seed_inds"):
    same(ep_r[k], ep_o[k], f"backbone {k}")
    checks += 1
ep_r["fp2_features"].sum().backward()
ep_o["fp2_features"].sum().backward()
grads_same(r_bb, o_bb, "backbone")
checks += 1

print(f"reference glue == oracle glue: {checks} bit-exact checks passed")
