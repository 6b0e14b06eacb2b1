"""GPU parity of the fused set-abstraction / feature-propagation path (libpn2_b200 GEMMs with fused
grouping / BatchNorm / ReLU / max-pool and the hand-written backward) against the CPU oracle modules
(the reference's Python glue restated over torch CPU fp32 conv / batch_norm / max_pool2d).
Index outputs bit-exact; float outputs within 1e-5 of the output's max magnitude (BASELINE.json
north_star: "within 1e-5 relative for the float feature paths"); gradients within 1e-4 in max-norm on the
small shapes.  At sizes with millions of ReLU / max-pool kinks a max-norm bound on gradients is ill-posed
(one flipped mask changes an element by O(1)); there the bound is not fitted to the observed error but
ARBITRATED against float64 (tests/arbiter.py): |ours - fp64| <= max(3 x |oracle_fp32 - fp64|, spec budget) + 2e-6 in
relative L2, where the spec budget is the gradient change a forward perturbation of the allowed 1e-5 causes in float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FEAT_TOL = 1e-5
GRAD_TOL = 1e-4
GEMM_TOL = 3e-6  # fp32 FFMA kernel: ~3e-7; tcgen05 3xTF32 split kernel: ~1e-6 (both vs float64)


@pytest.fixture(scope="module")
def K(built_lib):
    assert torch.cuda.is_available()
    import _pn2
    return _pn2


@pytest.fixture(scope="module")
def O():
    from oracle import pn2_oracle
    return pn2_oracle


@pytest.fixture(scope="module")
def A():
    import arbiter
    return arbiter


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- kernel-level checks --------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,k,n", [(1000, 20, 36), (128, 16, 128), (4097, 260, 256), (37, 4, 4), (70000, 132, 128)])
def test_gemm_plain_and_stats(rows, k, n, K):
    g = torch.Generator().manual_seed(rows)
    a = torch.randn(rows, k, generator=g).cuda()
    w = torch.randn(n, k, generator=g).cuda()
    wt, wp = K.mlp_prep_weights(w, 0, 0, k, n)
    assert torch.equal(wt, w.t().contiguous()) and torch.equal(wp, w)
    y, stats, tiles = K.mlp_forward(K.rows_plain(a, rows, k, k), k, n, wt, wp)
    want = a.double() @ w.double().t()
    assert rel(y, want) < GEMM_TOL
    assert rel(stats[:tiles, 0].double().sum(0), want.sum(0)) < 1e-5
    assert rel(stats[:tiles, 1].double().sum(0), (want * want).sum(0)) < 1e-5


def test_gemm_bnrelu_source_and_padding(K):
    g = torch.Generator().manual_seed(1)
    rows, c_in, c_out = 3000, 30, 10   # neither a multiple of 4
    kp, np_ = K.pad4(c_in), K.pad4(c_out)
    yprev = torch.zeros(rows, kp)
    yprev[:, :c_in] = torch.randn(rows, c_in, generator=g)
    scale, shift = torch.zeros(kp), torch.zeros(kp)
    scale[:c_in], shift[:c_in] = torch.randn(c_in, generator=g), torch.randn(c_in, generator=g)
    w = torch.randn(c_out, c_in, generator=g)
    wt, wp = K.mlp_prep_weights(w.cuda(), 0, 0, kp, np_)
    src = K.rows_bnrelu(yprev.cuda(), rows, kp, kp, scale.cuda(), shift.cuda())
    y, stats, tiles = K.mlp_forward(src, kp, np_, wt, wp)
    act = torch.relu(yprev[:, :c_in].double() * scale[:c_in].double() + shift[:c_in].double())
    want = act @ w.double().t()
    assert rel(y[:, :c_out], want) < GEMM_TOL
    assert float(y[:, c_out:].abs().max()) == 0.0


def test_gemm_gather_source_matches_query_and_group(K, O):
    xyz, feats = O.uniform_cloud(2, 1024, 5, seed=3)
    inds = O.ext.furthest_point_sampling(xyz, 128)
    new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    grouped, _, idx = O.query_and_group(0.2, 32, xyz, new_xyz, feats, use_xyz=True, normalize_xyz=True)
    w = torch.randn(24, 8, generator=torch.Generator().manual_seed(0))
    want = torch.einsum("oc,bcjs->bjso", w.double(), grouped.double()).reshape(-1, 24)
    feat_pm = K.to_point_major(feats.cuda())
    assert feat_pm.shape == (2 * 1024, 8) and float(feat_pm[:, 5:].abs().max()) == 0.0
    src = K.rows_gather(feat_pm, 8, 8, idx.cuda(), xyz.cuda(), new_xyz.cuda(), 1024, 128, 32, True, 0.2)
    wt, wp = K.mlp_prep_weights(w.cuda(), 1, 8, 12, 24)
    y, _, _ = K.mlp_forward(src, 12, 24, wt, wp)
    assert rel(y, want) < GEMM_TOL


def test_transposes_roundtrip(K):
    x = torch.randn(3, 37, 1001).cuda()
    pm = K.to_point_major(x)
    assert pm.shape == (3 * 1001, 40)
    assert torch.equal(pm.view(3, 1001, 40)[:, :, :37], x.transpose(1, 2))
    assert torch.equal(K.to_channel_major(pm, 3, 37, 1001), x)


# ---- module-level parity ----------------------------------------------------------------------------------
def _randomise_bn(mod, seed=5):
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data = torch.randn(m.weight.shape, generator=g)      # both signs: exercises max vs relu order
            m.bias.data = 0.3 * torch.randn(m.bias.shape, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)


def _pair(make_ours, make_oracle, seed=0, randomise_bn=False):
    torch.manual_seed(seed)
    ours = make_ours()
    oracle = make_oracle()
    if randomise_bn:
        _randomise_bn(ours)
    oracle.load_state_dict(ours.state_dict())
    return ours.cuda(), oracle


def _assert_feat(A, what, out, out_o, tol, run64):
    """Forward parity against the fp32 oracle; on disagreement float64 is the referee.  (Root cause of round 1's
    intermittent test_fp_matches_oracle[True] failure, found with this very message: ours was 2.8e-7 from float64 and
    bit-identical across 900 runs while torch CPU fp32's FIRST evaluation in a fresh process was 7.6e-5 off and differed
    from its own re-run -- profiles/r2_flake_hunt.txt.)"""
    r = rel(out, out_o)
    if r <= tol:
        return
    out_x = run64()
    e_ours, e_ref = A.rel_max(out, out_x), A.rel_max(out_o, out_x)
    assert e_ours <= tol, f"{what}: ours vs oracle {r:.3e}; ours vs fp64 {e_ours:.3e}; oracle_fp32 vs fp64 {e_ref:.3e}"
    print(f"{what}: oracle_fp32 off by {e_ref:.3e} from fp64 (ours {e_ours:.3e})")


def _check_module_grads(ours, oracle, tol=GRAD_TOL):
    for (n1, p1), (n2, p2) in zip(ours.named_parameters(), oracle.named_parameters()):
        assert n1 == n2
        assert p1.grad is not None, n1
        assert rel(p1.grad, p2.grad) <= tol, (n1, rel(p1.grad, p2.grad))


def _check_bn_buffers(ours, oracle):
    for (n1, b1), (n2, b2) in zip(ours.named_buffers(), oracle.named_buffers()):
        if b1.dtype.is_floating_point:
            assert rel(b1, b2) <= 1e-5, n1
        else:
            assert torch.equal(b1.cpu(), b2), n1


@pytest.mark.parametrize("train", [True, False])
def test_sa_config1_matches_oracle(train, K, O):
    """BASELINE.json configs[0]: 2 x 1024 points, SA(npoint=512, r=0.2, nsample=64, mlp=[3,64])."""
    import pointnet2_modules as M
    kw = dict(npoint=512, radius=0.2, nsample=64, use_xyz=True, normalize_xyz=True)
    ours, oracle = _pair(lambda: M.PointnetSAModuleVotes(mlp=[3, 64], **kw), lambda: O.OracleSAModuleVotes(mlp=[3, 64], **kw),
                         randomise_bn=True)
    ours.train(train)
    oracle.train(train)
    xyz, feats = O.uniform_cloud(2, 1024, 3, seed=0)
    f_dev, f_cpu = feats.cuda().requires_grad_(True), feats.clone().requires_grad_(True)
    before = K.launch_count
    new_xyz, out, inds = ours(xyz.cuda(), f_dev)
    new_xyz_o, out_o, inds_o = oracle(xyz, f_cpu)
    assert K.launch_count - before >= 6, "fused path did not run"
    assert inds.dtype == torch.int32 and torch.equal(inds.cpu(), inds_o)
    assert torch.equal(new_xyz.cpu(), new_xyz_o)
    assert out.shape == out_o.shape and rel(out, out_o) <= FEAT_TOL
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(1))
    (out * cot.cuda()).sum().backward()
    (out_o * cot).sum().backward()
    assert rel(f_dev.grad, f_cpu.grad) <= GRAD_TOL
    _check_module_grads(ours, oracle)
    _check_bn_buffers(ours, oracle)


def test_sa_three_layer_xyz_grad_matches_oracle(K, O):
    """vote_aggregation-like: deeper MLP, xyz itself requires grad (models/pq_transformer.py:159-166,220)."""
    import pointnet2_modules as M
    kw = dict(npoint=128, radius=0.3, nsample=16, use_xyz=True, normalize_xyz=True)
    ours, oracle = _pair(lambda: M.PointnetSAModuleVotes(mlp=[20, 32, 24, 40], **kw),
                         lambda: O.OracleSAModuleVotes(mlp=[20, 32, 24, 40], **kw), seed=2)
    ours.train()
    oracle.train()
    xyz, feats = O.uniform_cloud(3, 700, 20, seed=11)
    x_dev, x_cpu = xyz.cuda().requires_grad_(True), xyz.clone().requires_grad_(True)
    f_dev, f_cpu = feats.cuda().requires_grad_(True), feats.clone().requires_grad_(True)
    new_xyz, out, inds = ours(x_dev, f_dev)
    new_xyz_o, out_o, inds_o = oracle(x_cpu, f_cpu)
    assert torch.equal(inds.cpu(), inds_o) and rel(out, out_o) <= FEAT_TOL
    g = torch.Generator().manual_seed(3)
    cot, cot_xyz = torch.randn(out_o.shape, generator=g), torch.randn(new_xyz_o.shape, generator=g)
    ((out * cot.cuda()).sum() + (new_xyz * cot_xyz.cuda()).sum()).backward()
    ((out_o * cot).sum() + (new_xyz_o * cot_xyz).sum()).backward()
    assert rel(f_dev.grad, f_cpu.grad) <= GRAD_TOL
    assert rel(x_dev.grad, x_cpu.grad) <= GRAD_TOL
    _check_module_grads(ours, oracle)


@pytest.mark.parametrize("mlp,nsample", [([128, 48], 5), ([128, 40, 24], 5), ([256, 32, 16], 16), ([128, 36], 64)])
def test_sa_weight_gradient_xyz_block_layouts(mlp, nsample, K, O):
    """Weight gradient of the gathered first layer when the feature width is a multiple of 128, i.e. the 4-column xyz block
    starts a new 128-column tile: one layer = pooled dY source, xyz block in an n-tile of its own; several layers = plain dY
    source, xyz block folded into the last feature tile as MMA columns 128..143 (mlp_gemm_tc.cu, WgCfg::kFoldable).
    Odd nsample / position counts that are no multiple of 32: partial k-blocks, pool slots that wrap inside a k-block."""
    import pointnet2_modules as M
    kw = dict(npoint=37, radius=0.35, nsample=nsample, use_xyz=True, normalize_xyz=True)
    ours, oracle = _pair(lambda: M.PointnetSAModuleVotes(mlp=list(mlp), **kw), lambda: O.OracleSAModuleVotes(mlp=list(mlp), **kw),
                         seed=13, randomise_bn=True)
    ours.train()
    oracle.train()
    xyz, feats = O.uniform_cloud(2, 300, mlp[0], seed=21)
    x_dev, x_cpu = xyz.cuda().requires_grad_(True), xyz.clone().requires_grad_(True)
    f_dev, f_cpu = feats.cuda().requires_grad_(True), feats.clone().requires_grad_(True)
    new_xyz, out, inds = ours(x_dev, f_dev)
    new_xyz_o, out_o, inds_o = oracle(x_cpu, f_cpu)
    assert torch.equal(inds.cpu(), inds_o) and rel(out, out_o) <= FEAT_TOL
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(5))
    (out * cot.cuda()).sum().backward()
    (out_o * cot).sum().backward()
    assert rel(f_dev.grad, f_cpu.grad) <= GRAD_TOL
    assert rel(x_dev.grad, x_cpu.grad) <= GRAD_TOL
    _check_module_grads(ours, oracle)


def test_sa_without_features_and_given_inds(K, O):
    import pointnet2_modules as M
    kw = dict(npoint=64, radius=0.25, nsample=8, use_xyz=True, normalize_xyz=False)
    ours, oracle = _pair(lambda: M.PointnetSAModuleVotes(mlp=[0, 16, 16], **kw), lambda: O.OracleSAModuleVotes(mlp=[0, 16, 16], **kw))
    xyz, _ = O.uniform_cloud(2, 333, 3, seed=4)
    inds = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).integers(0, 333, (2, 64)).astype(np.int32))
    new_xyz, out, inds2 = ours(xyz.cuda(), None, inds.cuda())
    new_xyz_o, out_o, _ = oracle(xyz, None, inds)
    assert torch.equal(inds2.cpu(), inds) and torch.equal(new_xyz.cpu(), new_xyz_o)
    assert rel(out, out_o) <= FEAT_TOL
    out.sum().backward()
    out_o.sum().backward()
    _check_module_grads(ours, oracle)


@pytest.mark.parametrize("train", [True, False])
def test_fp_matches_oracle(train, K, O, A):
    import pointnet2_modules as M
    ours, oracle = _pair(lambda: M.PointnetFPModule(mlp=[64 + 12, 48, 20]), lambda: O.OracleFPModule(mlp=[64 + 12, 48, 20]), seed=7,
                         randomise_bn=True)
    ours.train(train)
    oracle.train(train)
    exact = A.to64(oracle)
    unknown, uf = O.uniform_cloud(2, 900, 12, seed=21)
    known, kf = O.uniform_cloud(2, 150, 64, seed=22)
    uf_d, kf_d = uf.cuda().requires_grad_(True), kf.cuda().requires_grad_(True)
    uf_c, kf_c = uf.clone().requires_grad_(True), kf.clone().requires_grad_(True)
    state_o = {k: v.clone() for k, v in oracle.state_dict().items()}
    out = ours(unknown.cuda(), known.cuda(), uf_d, kf_d)
    out_o = oracle(unknown, known, uf_c, kf_c)
    bad_oracle = rel(out, out_o) > FEAT_TOL
    _assert_feat(A, "FP out", out, out_o, FEAT_TOL, lambda: A.fp_forward64(exact, O, unknown, known, uf.double(), kf.double()))
    if bad_oracle:  # the fp32 oracle's first evaluation was the inaccurate one (see _assert_feat): evaluate it again
        oracle.load_state_dict(state_o)
        out_o = oracle(unknown, known, uf_c, kf_c)
        assert rel(out, out_o) <= FEAT_TOL
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(1))
    (out * cot.cuda()).sum().backward()
    (out_o * cot).sum().backward()
    assert rel(uf_d.grad, uf_c.grad) <= GRAD_TOL and rel(kf_d.grad, kf_c.grad) <= GRAD_TOL
    _check_module_grads(ours, oracle)
    _check_bn_buffers(ours, oracle)


def test_unfused_fallback_paths_still_work(K, O):
    """avg pooling is not fused: it must run on the op-level kernels and agree with the fused result's
    pre-pool activations in the mean (sanity), and MSG modules must run at all."""
    import pointnet2_modules as M
    torch.manual_seed(0)
    sa = M.PointnetSAModuleVotes(npoint=64, radius=0.3, nsample=16, mlp=[3, 8], pooling='avg', normalize_xyz=True).cuda()
    xyz, feats = O.uniform_cloud(2, 256, 3, seed=1)
    _, out, _ = sa(xyz.cuda(), feats.cuda())
    assert out.shape == (2, 8, 64) and torch.isfinite(out).all()
    msg = M.PointnetSAModuleMSG(npoint=32, radii=[0.2, 0.4], nsamples=[8, 16], mlps=[[3, 8], [3, 16]]).cuda()
    new_xyz, f = msg(xyz.cuda(), feats.cuda())
    assert f.shape == (2, 24, 32) and new_xyz.shape == (2, 32, 3)
    f.sum().backward()


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _named_grads(module, prefix=""):
    return {f"{prefix}grad {n}": p.grad.detach().clone() for n, p in module.named_parameters()}


def _backbone_pair(O):
    from backbone import Pointnet2Backbone
    torch.manual_seed(0)
    ours = Pointnet2Backbone(input_feature_dim=3)
    oracle = O.OracleBackbone(input_feature_dim=3)
    oracle.load_state_dict(ours.state_dict())
    return ours.cuda().train(), oracle.train()


@pytest.mark.parametrize("npts", [8192, 40000])
def test_backbone_chain_matches_oracle(npts, K, O, A):
    """BASELINE.json configs[1]: ScanNet-shaped 40000 x 6 cloud through the full backbone, fwd+bwd.

    Index outputs must be bit-exact through all four levels.  Features: sa1 within 1e-5 of the oracle; the deeper
    stages are a CHAIN of 18 training-mode BatchNorm layers in which each implementation feeds on its own slightly
    different outputs (torch CPU fp32 itself is 4e-6 ... 6e-6 away from float64 at sa4 / fp2), so they and every
    parameter gradient are arbitrated against the float64 evaluation of the same chain (tests/arbiter.py):
    |ours - fp64| <= max(3 x |oracle_fp32 - fp64|, spec budget) + 2e-6, the budget being the change a forward
    perturbation of the allowed 1e-5 causes in float64.  test_backbone_stages_identical_inputs holds every module to
    1e-5 on identical inputs."""
    ours, oracle = _backbone_pair(O)
    exact = A.to64(oracle)
    cloud = O.scannet_like_cloud(npts, seed=1234)[None]
    ep = ours(cloud.cuda())
    ep_o = oracle(cloud)
    for k in ["sa1_inds", "sa2_inds", "fp2_inds"]:
        assert torch.equal(ep[k].cpu(), ep_o[k]), k
    for k in ["sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"]:
        assert torch.equal(ep[k].cpu(), ep_o[k]), k
    assert rel(ep["sa1_features"], ep_o["sa1_features"]) <= FEAT_TOL
    feats = ["sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"]
    cot = torch.randn(ep_o["fp2_features"].shape, generator=torch.Generator().manual_seed(1))
    (ep["fp2_features"] * cot.cuda()).sum().backward()
    (ep_o["fp2_features"] * cot).sum().backward()

    def run64():
        exact.zero_grad()
        ep_x = A.backbone_forward64(exact, O, cloud)
        (ep_x["fp2_features"] * cot.double()).sum().backward()
        return {**{k: ep_x[k].detach().clone() for k in feats}, **_named_grads(exact)}

    print(A.arbitrate(f"backbone[{npts}]", {**{k: ep[k] for k in feats}, **_named_grads(ours)},
                      {**{k: ep_o[k] for k in feats}, **_named_grads(oracle)}, run64, exact))
    _check_bn_buffers(ours, oracle)


def test_backbone_stages_identical_inputs(K, O, A):
    """Every SA / FP module of the 40k-point backbone, fed the ORACLE's inputs for that stage (identical
    inputs on both sides): indices bit-exact, features within 1e-5 of the oracle, gradients arbitrated against
    float64 (33M ReLU kinks per layer make a max-norm bound ill-posed; the small-shape tests above hold gradients to
    1e-4 in max-norm)."""
    ours, oracle = _backbone_pair(O)
    exact = A.to64(oracle)
    cloud = O.scannet_like_cloud(40000, seed=1234)[None]
    with torch.no_grad():
        ep_o = oracle(cloud)
    xyz0 = cloud[..., :3].contiguous()
    f0 = cloud[..., 3:].transpose(1, 2).contiguous()
    stages = [("sa1", xyz0, f0), ("sa2", ep_o["sa1_xyz"], ep_o["sa1_features"]),
              ("sa3", ep_o["sa2_xyz"], ep_o["sa2_features"]), ("sa4", ep_o["sa3_xyz"], ep_o["sa3_features"])]
    for name, xyz, feats in stages:
        m_d, m_c, m_x = getattr(ours, name), getattr(oracle, name), getattr(exact, name)
        f_d, f_c = feats.cuda().requires_grad_(True), feats.clone().requires_grad_(True)
        nx, out, inds = m_d(xyz.cuda(), f_d)
        nx_o, out_o, inds_o = m_c(xyz, f_c)
        assert torch.equal(inds.cpu(), inds_o) and torch.equal(nx.cpu(), nx_o), name
        assert rel(out, out_o) <= FEAT_TOL, (name, rel(out, out_o))
        cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(2))
        (out * cot.cuda()).sum().backward()
        (out_o * cot).sum().backward()

        def run64():
            m_x.zero_grad()
            f_x = feats.double().requires_grad_(True)
            _, out_x, _ = A.sa_forward64(m_x, O, xyz, f_x)
            (out_x * cot.double()).sum().backward()
            return {"d features": f_x.grad.clone(), **_named_grads(m_x)}

        print(A.arbitrate(name, {"d features": f_d.grad, **_named_grads(m_d)}, {"d features": f_c.grad, **_named_grads(m_c)},
                          run64, m_x))
    with torch.no_grad():
        fp1_o = oracle.fp1(ep_o["sa3_xyz"], ep_o["sa4_xyz"], ep_o["sa3_features"], ep_o["sa4_features"])
    for name, args in [("fp1", (ep_o["sa3_xyz"], ep_o["sa4_xyz"], ep_o["sa3_features"], ep_o["sa4_features"])),
                       ("fp2", (ep_o["sa2_xyz"], ep_o["sa3_xyz"], ep_o["sa2_features"], fp1_o))]:
        m_d, m_c, m_x = getattr(ours, name), getattr(oracle, name), getattr(exact, name)
        a_d = [t.cuda().requires_grad_(i >= 2) for i, t in enumerate(args)]
        a_c = [t.clone().requires_grad_(i >= 2) for i, t in enumerate(args)]
        out, out_o = m_d(*a_d), m_c(*a_c)
        assert rel(out, out_o) <= FEAT_TOL, (name, rel(out, out_o))
        cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(3))
        (out * cot.cuda()).sum().backward()
        (out_o * cot).sum().backward()

        def run64():
            m_x.zero_grad()
            a_x = [t.double().requires_grad_(True) for t in args[2:]]
            out_x = A.fp_forward64(m_x, O, args[0], args[1], a_x[0], a_x[1])
            (out_x * cot.double()).sum().backward()
            return {"d unknow_feats": a_x[0].grad.clone(), "d known_feats": a_x[1].grad.clone(), **_named_grads(m_x)}

        print(A.arbitrate(name, {"d unknow_feats": a_d[2].grad, "d known_feats": a_d[3].grad, **_named_grads(m_d)},
                          {"d unknow_feats": a_c[2].grad, "d known_feats": a_c[3].grad, **_named_grads(m_c)}, run64, m_x))


def test_ffma_kernel_path_still_green():
    """The tcgen05 (3xTF32) GEMM is the default; PN2_TC=0 selects the fp32 FFMA kernel everywhere.  Re-run the
    kernel-level and module-level checks of this file on that path in a fresh process (no retry: round 1's
    intermittent failure of test_fp_matches_oracle[True] was hunted with tools/flake_hunt.py -- 900 in-process
    repetitions on both paths, bit-identical results from run to run, profiles/r2_flake_hunt.txt)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PN2_TC="0")
    cmd = [sys.executable, "-m", "pytest", __file__, "-m", "gpu", "-q", "-x", "-k",
           "gemm or transposes or config1 or three_layer or without_features or fp_matches"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-4000:]


def test_config3_callers_vote_aggregation_and_fps_module_batch8(K, O, A):
    """BASELINE.json configs[2] exercises the path through the detector's callers: FPSModule
    (models/utils/pointnet_util.py:52-69 = furthest_point_sample + two gather_operations) and
    vote_aggregation = PointnetSAModuleVotes(256, 0.3, 16, [288+3,288,288,288]) (models/pq_transformer.py:159-166)
    on a batch of 8 clouds of 1024 seeds, eval mode (BN running statistics) and train mode, xyz carrying gradients.
    (The real pq_transformer.py / pointnet_util.py files run in tests/test_gpu_real_callers.py.)"""
    import pointnet2_modules as M
    import pointnet2_utils as U
    kw = dict(npoint=256, radius=0.3, nsample=16, use_xyz=True, normalize_xyz=True)
    ours, oracle = _pair(lambda: M.PointnetSAModuleVotes(mlp=[288, 288, 288, 288], **kw),
                         lambda: O.OracleSAModuleVotes(mlp=[288, 288, 288, 288], **kw), seed=3, randomise_bn=True)
    exact = A.to64(oracle)
    xyz, feats = O.uniform_cloud(8, 1024, 288, seed=31)
    xyz = xyz * 0.9
    # FPSModule
    inds = U.furthest_point_sample(xyz.cuda(), 256)
    inds_o = O.furthest_point_sample(xyz, 256)
    assert torch.equal(inds.cpu(), inds_o)
    g_xyz = U.gather_operation(xyz.cuda().transpose(1, 2).contiguous(), inds)
    g_feat = U.gather_operation(feats.cuda(), inds)
    assert torch.equal(g_xyz.cpu(), O.gather_operation(xyz.transpose(1, 2).contiguous(), inds_o))
    assert torch.equal(g_feat.cpu(), O.gather_operation(feats, inds_o))
    for train in (False, True):
        for m in (ours, oracle, exact):
            m.train(train)
        x_d, x_c = xyz.cuda().requires_grad_(True), xyz.clone().requires_grad_(True)
        f_d, f_c = feats.cuda().requires_grad_(True), feats.clone().requires_grad_(True)
        nx, out, ii = ours(x_d, f_d)
        nx_o, out_o, ii_o = oracle(x_c, f_c)
        assert torch.equal(ii.cpu(), ii_o) and torch.equal(nx.cpu(), nx_o)
        assert rel(out, out_o) <= (FEAT_TOL if not train else 2 * FEAT_TOL), (train, rel(out, out_o))
        cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(5))
        (out * cot.cuda()).sum().backward()
        (out_o * cot).sum().backward()

        def run64():
            exact.zero_grad()
            x_x, f_x = xyz.double().requires_grad_(True), feats.double().requires_grad_(True)
            _, out_x, _ = A.sa_forward64(exact, O, xyz, f_x, xyz64=x_x)
            (out_x * cot.double()).sum().backward()
            return {"d features": f_x.grad.clone(), "d xyz": x_x.grad.clone(), **_named_grads(exact)}

        print(A.arbitrate(f"vote_aggregation train={train}", {"d features": f_d.grad, "d xyz": x_d.grad, **_named_grads(ours)},
                          {"d features": f_c.grad, "d xyz": x_c.grad, **_named_grads(oracle)}, run64, exact))
        for m in (ours, oracle, exact):
            m.zero_grad()


def test_config5_arkit_fp_stress(K, O, A):
    """BASELINE.json configs[4]: ARKit-shaped 50000-point cloud (centred, yawed: points near the origin exercise
    the FPS skip), PointnetFPModule(mlp=[256+3,256,128]) from SA1's 2048 points to all 50000, fwd+bwd."""
    import pointnet2_modules as M
    cloud = O.scannet_like_cloud(50000, seed=4321, centred=True, yaw=True)
    xyz = cloud[None, :, :3].contiguous()
    col = cloud[None, :, 3:].transpose(1, 2).contiguous()
    inds_o = O.furthest_point_sample(xyz, 2048)
    inds = K.furthest_point_sampling(xyz.cuda(), 2048)
    assert torch.equal(inds.cpu(), inds_o)
    known = torch.gather(xyz, 1, inds_o.long()[..., None].expand(-1, -1, 3)).contiguous()
    kf = torch.randn(1, 256, 2048, generator=torch.Generator().manual_seed(8))
    ours, oracle = _pair(lambda: M.PointnetFPModule(mlp=[256 + 3, 256, 128]), lambda: O.OracleFPModule(mlp=[256 + 3, 256, 128]), seed=9)
    exact = A.to64(oracle)
    for m in (ours, oracle, exact):
        m.train()
    c_d, k_d = col.cuda().requires_grad_(True), kf.cuda().requires_grad_(True)
    c_c, k_c = col.clone().requires_grad_(True), kf.clone().requires_grad_(True)
    out = ours(xyz.cuda(), known.cuda(), c_d, k_d)
    out_o = oracle(xyz, known, c_c, k_c)
    assert rel(out, out_o) <= FEAT_TOL, rel(out, out_o)
    cot = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(6))
    (out * cot.cuda()).sum().backward()
    (out_o * cot).sum().backward()

    def run64():
        exact.zero_grad()
        c_x, k_x = col.double().requires_grad_(True), kf.double().requires_grad_(True)
        out_x = A.fp_forward64(exact, O, xyz, known, c_x, k_x)
        (out_x * cot.double()).sum().backward()
        return {"d unknow_feats": c_x.grad.clone(), "d known_feats": k_x.grad.clone(), **_named_grads(exact)}

    print(A.arbitrate("fp stress", {"d unknow_feats": c_d.grad, "d known_feats": k_d.grad, **_named_grads(ours)},
                      {"d unknow_feats": c_c.grad, "d known_feats": k_c.grad, **_named_grads(oracle)}, run64, exact))


# ---- fused three_nn: index path under ties ------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["lattice", "duplicates", "few_known"])
def test_fused_fp_three_nn_indices_bit_exact_under_ties(case, K, O):
    """fp_interpolate_kernel re-implements three_nn (interpolate_gpu.cu:14-64) inside the fused FP front end; its
    index output must equal the reference kernel's on inputs FULL of exact distance ties: known points on an integer
    lattice (many equidistant neighbours), duplicated known points (distance ties at every rank, ties -> lower index),
    and m < 3 known points (unfilled slots keep index 0)."""
    g = np.random.Generator(np.random.PCG64(17))
    if case == "lattice":
        known = torch.from_numpy(np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(1, -1, 3).astype(np.float32))
        known = known[:, torch.from_numpy(g.permutation(known.shape[1]))].repeat(2, 1, 1).contiguous()
        unknown = torch.from_numpy(g.integers(0, 11, (2, 700, 3)).astype(np.float32) * 0.5)   # lattice points and cell centres
    elif case == "duplicates":
        base = torch.from_numpy(g.random((2, 50, 3)).astype(np.float32))
        known = torch.cat([base, base, base[:, :17]], dim=1).contiguous()
        unknown = torch.cat([torch.from_numpy(g.random((2, 300, 3)).astype(np.float32)), base], dim=1).contiguous()
    else:
        known = torch.from_numpy(g.random((3, 2, 3)).astype(np.float32))
        unknown = torch.from_numpy(g.random((3, 40, 3)).astype(np.float32))
    b, n, m = unknown.shape[0], unknown.shape[1], known.shape[1]
    c = 8
    kf = torch.from_numpy(g.standard_normal((b, c, m)).astype(np.float32))
    known_pm = K.to_point_major(kf.cuda())
    x = torch.empty(b * n, c, device="cuda")
    idx, w = K.fp_interpolate(unknown.cuda(), known.cuda(), known_pm, c, x, c)
    d2_o, idx_o = O.ext.three_nn(unknown, known)
    assert torch.equal(idx.cpu(), idx_o), case
    if m >= 3:
        w_o = O.fp_weights(torch.sqrt(d2_o))
        assert torch.allclose(w.cpu(), w_o, rtol=1e-6, atol=1e-12), case
        want = O.ext.three_interpolate(kf, idx_o, w_o).transpose(1, 2).reshape(b * n, c)
        assert torch.allclose(x.cpu(), want, rtol=1e-5, atol=1e-6), case


# ---- the timed path: whole-step CUDA graph == eager ----------------------------------------------------------------
@pytest.mark.parametrize("prefetch", [False, True])
def test_graphed_train_step_matches_eager(prefetch, K, O):
    """bench.py times graphed.GraphedTrainStep replays (forward + backward + the side-stream fork/join in ONE CUDA
    graph, gradients produced in the arena); every other parity test runs eagerly.  Over 4 replays with rotating
    inputs the replay must reproduce the eager step: the forward output and the BatchNorm buffers bit for bit (no atomics
    in the forward), every gradient to 2e-5 of its max (the scatter-add backward kernels use fp32 atomics like the
    reference's group_points_grad / three_interpolate_grad, so the last bits depend on the launch and are then amplified
    by the BatchNorm backward sums upstream: measured 3e-6 ... 1e-5 between two eager runs as well)."""
    from backbone import Pointnet2Backbone
    from graphed import GraphedTrainStep
    torch.manual_seed(0)
    model = Pointnet2Backbone(input_feature_dim=3).cuda().train()

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = model

        def forward(self, cloud):
            return self.backbone(cloud)["fp2_features"]

    net = Net()
    scenes = torch.stack([O.scannet_like_cloud(20000, seed=77 + i) for i in range(3)]).cuda()
    cot = torch.randn(1, 288, 1024, generator=torch.Generator().manual_seed(4)).cuda()
    outs = {}

    def loss_fn(out):
        outs["last"] = out
        return (out * cot).sum()

    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    # prefetch: the first level's FPS + ball query of the NEXT batch run on a side stream underneath the current step
    step = GraphedTrainStep(net, loss_fn, (scenes[0][None],),
                            prefetch=(model.sa1, lambda inp: inp[0][..., :3]) if prefetch else None)
    static_out = outs["last"]
    assert step.launches_per_step and step.launches_per_step > 50
    params = [p for p in model.parameters()]
    for it in range(4):
        cloud = scenes[it % 3][None]
        model.load_state_dict(state0)           # same BatchNorm running statistics on both sides
        loss_g = (step(cloud, next_inputs=(scenes[(it + 1) % 3][None],)) if prefetch else step(cloud)).clone()
        out_g = static_out.clone()
        grads_g = [p.grad.clone() for p in params]
        bufs_g = {k: v.clone() for k, v in model.state_dict().items()}
        model.load_state_dict(state0)
        model.zero_grad(set_to_none=True)
        out_e = net(cloud)
        loss_e = (out_e * cot).sum()
        loss_e.backward()
        assert torch.equal(out_g, out_e), it
        assert torch.equal(loss_g, loss_e.detach()), it
        for (n, p), g in zip(model.named_parameters(), grads_g):
            d = float((g - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-30))
            assert d <= 2e-5, (it, n, d)
        for k, v in model.state_dict().items():
            assert torch.equal(v, bufs_g[k]), (it, k)
