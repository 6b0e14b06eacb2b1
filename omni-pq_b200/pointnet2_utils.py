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

Layout of this file: the two op families are described once (`_index_op`: outputs are indices, nothing
flows back; `_indexed_copy_op`: a gather whose backward is the matching scatter-add kernel) and the six
public classes are instances of those descriptions; tests/test_utils_glue_cpu.py runs every one of them,
forward and backward, on the CPU oracle's kernels.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

import _pn2 as _ext
import pytorch_utils as pt_utils  # noqa: F401  (reference modules reach pt_utils through this name)


def _index_op(name, n_inputs, run, doc):
    """autograd.Function whose result is an int32 index tensor (plus optional float companions): the indices are
    marked non-differentiable and every input receives None."""

    def forward(ctx, *args):
        out = run(*args)
        ctx.mark_non_differentiable(out[-1] if isinstance(out, tuple) else out)
        return out

    def backward(ctx, *grads):
        return (None,) * n_inputs

    return type(name, (Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward), "__doc__": doc})


def _indexed_copy_op(name, n_inputs, fwd_kernel, bwd_kernel, src_extent_dim, doc):
    """autograd.Function `out = kernel(src, index, *rest)` that copies (or blends) columns of `src` selected by an
    index tensor; the gradient w.r.t. `src` is the scatter-add kernel, which needs the index operands and the
    extent of the source's last axis.  Only `src` is differentiable."""

    def forward(ctx, src, *index_operands):
        ctx.pn2_saved = (index_operands, src.size(src_extent_dim))
        return getattr(_ext, fwd_kernel)(src, *index_operands)

    def backward(ctx, upstream):
        index_operands, extent = ctx.pn2_saved
        d_src = getattr(_ext, bwd_kernel)(upstream.contiguous(), *index_operands, extent)
        return (d_src,) + (None,) * (n_inputs - 1)

    return type(name, (Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward), "__doc__": doc})


FurthestPointSampling = _index_op(
    "FurthestPointSampling", 2, lambda xyz, npoint: _ext.furthest_point_sampling(xyz, npoint),
    "(xyz (B,N,3) f32, npoint) -> (B,npoint) int32; starts at index 0, reference tie-break order.")
furthest_point_sample = FurthestPointSampling.apply

GatherOperation = _indexed_copy_op(
    "GatherOperation", 2, "gather_points", "gather_points_grad", 2,
    "(features (B,C,N), idx (B,npoint) int32) -> (B,C,npoint).")
gather_operation = GatherOperation.apply


def _three_nn(unknown, known):
    squared, which = _ext.three_nn(unknown, known)
    return torch.sqrt(squared), which  # the kernel reports squared distances (pointnet2_utils.py:140-142)


ThreeNN = _index_op(
    "ThreeNN", 2, _three_nn,
    "(unknown (B,n,3), known (B,m,3)) -> (l2 distances (B,n,3) ascending, idx (B,n,3) int32).")
three_nn = ThreeNN.apply

ThreeInterpolate = _indexed_copy_op(
    "ThreeInterpolate", 3, "three_interpolate", "three_interpolate_grad", 2,
    "(features (B,c,m), idx (B,n,3) int32, weight (B,n,3)) -> (B,c,n), the 3-tap weighted sum.")
three_interpolate = ThreeInterpolate.apply

GroupingOperation = _indexed_copy_op(
    "GroupingOperation", 2, "group_points", "group_points_grad", 2,
    "(features (B,C,N), idx (B,npoint,nsample) int32) -> (B,C,npoint,nsample).")
grouping_operation = GroupingOperation.apply

# Python argument order (radius, nsample, xyz, new_xyz); the kernel takes the centres first (:282)
BallQuery = _index_op(
    "BallQuery", 4, lambda radius, nsample, xyz, new_xyz: _ext.ball_query(new_xyz, xyz, radius, nsample),
    "(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32.")
ball_query = BallQuery.apply


class RandomDropout(nn.Module):
    """Import compatibility only: like the reference's (:40-48) it calls a helper `pytorch_utils` never had, so
    using it raises AttributeError there and here."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p, self.inplace = p, inplace

    def forward(self, X):
        theta = torch.Tensor(1).uniform_(0, self.p)[0]
        return pt_utils.feature_dropout_no_scaling(X, theta, self.train, self.inplace)


def _stack_channels(local_xyz, grouped_features, use_xyz):
    """Channel layout every grouper produces: local coordinates first, then the features."""
    if grouped_features is None:
        assert use_xyz, "Cannot have not features and not use xyz as a feature!"
        return local_xyz
    return torch.cat([local_xyz, grouped_features], dim=1) if use_xyz else grouped_features


class QueryAndGroup(nn.Module):
    """(xyz, new_xyz, features) -> (B, 3+C, npoint, nsample): ball query around every centre, neighbours'
    coordinates relative to the centre (divided by the radius when `normalize_xyz`) stacked in front of their
    features; optionally also the local coordinates and the per-ball unique counts."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        assert sample_uniformly or not ret_unique_cnt
        self.radius, self.nsample = radius, nsample
        self.use_xyz, self.normalize_xyz = use_xyz, normalize_xyz
        self.ret_grouped_xyz, self.ret_unique_cnt = ret_grouped_xyz, ret_unique_cnt
        self.sample_uniformly = sample_uniformly

    def _redraw_padding(self, idx):
        """:336-345 -- a ball with k < nsample members is padded with its first hit by the kernel; replace that
        padding by uniform re-draws of the k members (host loop; no PQ-Transformer config enables it)."""
        counts = torch.zeros(idx.shape[:2])
        for cloud, centre in ((c, j) for c in range(idx.shape[0]) for j in range(idx.shape[1])):
            members = torch.unique(idx[cloud, centre])
            k = members.numel()
            counts[cloud, centre] = k
            refill = members[torch.randint(0, k, (self.nsample - k,), dtype=torch.long)]
            idx[cloud, centre] = torch.cat((members, refill))
        return counts

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        counts = self._redraw_padding(idx) if self.sample_uniformly else None
        local = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        local -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            local /= self.radius
        grouped = grouping_operation(features, idx) if features is not None else None
        result = (_stack_channels(local, grouped, self.use_xyz),)
        result += (local,) if self.ret_grouped_xyz else ()
        result += (counts,) if self.ret_unique_cnt else ()
        return result if len(result) > 1 else result[0]


class GroupAll(nn.Module):
    """One group holding the whole cloud: -> (B, 3+C, 1, N)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        # the reference drops ret_grouped_xyz in __init__ (:387-390) and then reads it in forward, which raises;
        # the attribute is kept here so the documented behaviour works
        self.use_xyz, self.ret_grouped_xyz = use_xyz, ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        everything = xyz.transpose(1, 2).unsqueeze(2)
        stacked = everything if features is None else _stack_channels(everything, features.unsqueeze(2), self.use_xyz)
        return (stacked, everything) if self.ret_grouped_xyz else stacked
