"""The Python glue of `pointnet2_utils.py` (argument order, saved state, gradients, channel layout of the groupers),
run on the CPU: the module's `_ext` binding is replaced by the oracle's CPU kernels (`oracle.pn2_oracle.ext`, the same
function names and signatures as the reference's `_ext`) and every public Function / Module is compared, forward and
backward, with the oracle's own restatement of the reference glue.  The GPU suite checks the same surface on the
real kernels; this keeps the host logic honest on machines without a GPU."""
import importlib

import pytest
import torch

from oracle import pn2_oracle as O


@pytest.fixture()
def U(built_lib, monkeypatch):
    mod = importlib.import_module("pointnet2_utils")
    monkeypatch.setattr(mod, "_ext", O.ext)
    return mod


def _cloud(b=2, n=96, c=5, seed=0):
    xyz, feats = O.uniform_cloud(b, n, c, seed=seed)
    return xyz, feats


def test_index_functions_match_oracle_and_are_non_differentiable(U):
    xyz, _ = _cloud()
    xyz = xyz.clone().requires_grad_(True)
    inds = U.furthest_point_sample(xyz, 24)
    assert inds.dtype == torch.int32 and not inds.requires_grad
    assert torch.equal(inds, O.furthest_point_sample(xyz.detach(), 24))
    centres = torch.gather(xyz.detach(), 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = U.ball_query(0.35, 8, xyz, centres)
    assert idx.shape == (2, 24, 8) and not idx.requires_grad
    assert torch.equal(idx, O.ball_query(0.35, 8, xyz.detach(), centres))
    dist, nn_idx = U.three_nn(xyz, centres)
    dist_o, nn_o = O.three_nn(xyz.detach(), centres)
    assert torch.equal(nn_idx, nn_o) and torch.equal(dist, dist_o) and not nn_idx.requires_grad


def test_gather_group_interpolate_forward_and_backward_match_oracle(U):
    xyz, feats = _cloud(seed=3)
    inds = O.furthest_point_sample(xyz, 20)
    centres = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = O.ball_query(0.4, 6, xyz, centres)
    dist, nn_idx = O.three_nn(xyz, centres)
    w = O.fp_weights(dist)
    known = torch.randn(2, 7, 20, generator=torch.Generator().manual_seed(1))
    for ours, theirs, args in ((U.gather_operation, O.gather_operation, (feats, inds)),
                               (U.grouping_operation, O.grouping_operation, (feats, idx)),
                               (U.three_interpolate, O.three_interpolate, (known, nn_idx, w))):
        a = args[0].clone().requires_grad_(True)
        b = args[0].clone().requires_grad_(True)
        out, ref = ours(a, *args[1:]), theirs(b, *args[1:])
        assert torch.equal(out, ref)
        cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
        (out * cot).sum().backward()
        (ref * cot).sum().backward()
        assert torch.allclose(a.grad, b.grad, rtol=0, atol=1e-6)


@pytest.mark.parametrize("use_xyz,normalize,with_features", [(True, True, True), (True, False, True), (False, False, True),
                                                              (True, True, False)])
def test_query_and_group_matches_oracle(U, use_xyz, normalize, with_features):
    xyz, feats = _cloud(seed=5)
    inds = O.furthest_point_sample(xyz, 16)
    centres = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    f1 = feats.clone().requires_grad_(True) if with_features else None
    f2 = feats.clone().requires_grad_(True) if with_features else None
    grouper = U.QueryAndGroup(0.4, 8, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize)
    out, local = grouper(xyz, centres, f1)
    ref = O.query_and_group(0.4, 8, xyz, centres, f2, use_xyz=use_xyz, normalize_xyz=normalize)
    ref_out = ref[0] if isinstance(ref, (tuple, list)) else ref
    assert out.shape == ref_out.shape and torch.equal(out, ref_out)
    assert local.shape == (2, 3, 16, 8)
    if with_features:
        out.square().sum().backward()
        ref_out.square().sum().backward()
        assert torch.allclose(f1.grad, f2.grad, rtol=0, atol=1e-6)
    plain = U.QueryAndGroup(0.4, 8, use_xyz=use_xyz, normalize_xyz=normalize)(xyz, centres, feats if with_features else None)
    assert torch.is_tensor(plain) and torch.equal(plain, out.detach())


def test_group_all_and_uniform_resampling(U):
    xyz, feats = _cloud(seed=7)
    ga = U.GroupAll(use_xyz=True, ret_grouped_xyz=True)
    stacked, everything = ga(xyz, None, feats)
    assert stacked.shape == (2, 3 + 5, 1, 96) and torch.equal(stacked[:, :3], everything)
    assert torch.equal(stacked[:, 3:, 0], feats) and torch.equal(everything[:, :, 0], xyz.transpose(1, 2))
    assert torch.equal(U.GroupAll(use_xyz=False)(xyz, None, feats), feats.unsqueeze(2))
    assert torch.equal(U.GroupAll()(xyz, None, None), everything)
    inds = O.furthest_point_sample(xyz, 8)
    centres = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    torch.manual_seed(0)
    out, counts = U.QueryAndGroup(0.15, 12, sample_uniformly=True, ret_unique_cnt=True)(xyz, centres, feats)
    idx = O.ball_query(0.15, 12, xyz, centres)
    assert out.shape == (2, 8, 8, 12) and counts.shape == (2, 8)
    for b in range(2):
        for j in range(8):
            assert counts[b, j] == torch.unique(idx[b, j]).numel()
    with pytest.raises(AssertionError):
        U.QueryAndGroup(0.1, 4, ret_unique_cnt=True)
