"""Autograd surface of the PointNet++ ops, backed by libpn2_b200.so (sm_100a).

Drop-in for the reference's `pointnet2/pointnet2_utils.py`: the same six `autograd.Function`s and
their `.apply` aliases (FurthestPointSampling :51-80, GatherOperation :83-117, ThreeNN :120-149,
ThreeInterpolate :152-206, GroupingOperation :209-257, BallQuery :260-291), `QueryAndGroup`
(:294-376), `GroupAll` (:379-425) and `RandomDropout` (:40-48), with identical argument order, int32
index outputs and gradient behaviour (index outputs are non-differentiable; float inputs receive
scatter-added gradients).  Put this directory ahead of the reference's `pointnet2/` on `sys.path` and
`models/backbone_module.py`, `models/pq_transformer.py`, `models/utils/pointnet_util.py` import it
unchanged.

`_ext` here is the ctypes binding (`_pn2.py`) of the C-ABI library; it raises ImportError if the
library has not been built -- there is no CPU or eager-PyTorch fallback.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

import _pn2 as _ext
import pytorch_utils as pt_utils  # noqa: F401  (reference modules reach pt_utils through this name)


class RandomDropout(nn.Module):
    """Kept for import compatibility; like the reference's it relies on a helper that
    `pytorch_utils` never defined (pointnet2_utils.py:48), so calling it raises AttributeError."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p, self.inplace = p, inplace

    def forward(self, X):
        theta = torch.Tensor(1).uniform_(0, self.p)[0]
        return pt_utils.feature_dropout_no_scaling(X, theta, self.train, self.inplace)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B,N,3) f32 -> (B,npoint) int32, start index 0, reference tie-break order."""
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, _, n = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) l2 distances ascending, idx (B,n,3) int32)."""
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B,c,m), idx/weight (B,n,3) -> (B,c,n)."""
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample)."""
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32."""
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping: (xyz, new_xyz, features) -> (B, 3+C, npoint, nsample) with the
    centre-relative (optionally radius-normalised) coordinates as the first three channels."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def _resample_uniformly(self, idx):
        """pointnet2_utils.py:336-345: replace the 'first hit' padding of every ball by random
        re-draws of its unique members (host-side; unused by the PQ-Transformer configs)."""
        unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
        for b in range(idx.shape[0]):
            for r in range(idx.shape[1]):
                members = torch.unique(idx[b, r, :])
                k = members.shape[0]
                unique_cnt[b, r] = k
                extra = torch.randint(0, k, (self.nsample - k,), dtype=torch.long)
                idx[b, r, :] = torch.cat((members, members[extra]))
        return unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None

        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz /= self.radius

        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz

        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Single group containing every point: -> (B, 3+C, 1, N)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        # the reference never stores ret_grouped_xyz (pointnet2_utils.py:387-390) and then reads it in
        # forward, which raises; keep the attribute so the documented behaviour works
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features
