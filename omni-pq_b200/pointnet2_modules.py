"""PointNet++ set-abstraction / feature-propagation modules on the B200 kernels.

Drop-in for the reference's `pointnet2/pointnet2_modules.py`: `PointnetSAModuleVotes` (:164-272) and
`PointnetFPModule` (:356-416) are the two classes `models/backbone_module.py:18` and
`models/pq_transformer.py:14` import; `_PointnetSAModuleBase`, `PointnetSAModuleMSG`,
`PointnetSAModule`, `PointnetSAModuleMSGVotes` and `PointnetLFPModuleMSG` keep their names,
constructor arguments, attribute names (`mlp_module`, `mlp`, `mlps`, `groupers`, `grouper`) and
therefore their `state_dict` keys.

Two execution paths, both on libpn2_b200.so kernels only:

* fused (`fused.py`): `PointnetSAModuleVotes` with max pooling and `PointnetFPModule` run as a short
  chain of fused kernels -- FPS (+centre gather), ball query (it writes the int32 index tensor, which
  the backward needs), a first MLP layer that gathers the neighbour rows by index while it stages its
  operand (the grouped tensor of the reference is never materialised), BatchNorm(+ReLU) folded into
  the next layer's operand load, BatchNorm + ReLU + max-pool of the last layer as one kernel over its
  (materialised) pre-activations, three_nn + interpolation as one gather-MAC kernel -- with a
  hand-written backward.  The `nn.Conv2d` / `nn.BatchNorm2d` (or `SyncBatchNorm`) children are only
  read as parameter holders there.
* op-level: every other configuration (MSG variants, avg / rbf pooling, sample_uniformly, GroupAll,
  CPU-less odd shapes) composes the nine op kernels through `pointnet2_utils` exactly like the
  reference does and calls the `SharedMLP` children.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import pointnet2_utils  # noqa: E402
import pytorch_utils as pt_utils  # noqa: E402
import fused  # noqa: E402


def _sample_centres(xyz, npoint, inds=None):
    """FPS (unless `inds` is given) + gather of the sampled coordinates -> (new_xyz (B,npoint,3), inds)."""
    if npoint is None:
        return None, inds
    if inds is None:
        inds = pointnet2_utils.furthest_point_sample(xyz, npoint)
    new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), inds)
    return new_xyz.transpose(1, 2).contiguous(), inds


def _max_over_samples(x):
    """(B,C,npoint,nsample) -> (B,C,npoint)."""
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


def _build_scales(module, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly):
    """Shared constructor body of the multi-scale modules: fills `groupers` and `mlps`.
    Like the reference it bumps `mlp_spec[0]` by 3 *in the caller's list* when use_xyz."""
    assert len(radii) == len(nsamples) == len(mlps)
    module.npoint = npoint
    module.groupers = nn.ModuleList()
    module.mlps = nn.ModuleList()
    for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
        if npoint is not None:
            grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz,
                                                    sample_uniformly=sample_uniformly)
        else:
            grouper = pointnet2_utils.GroupAll(use_xyz)
        module.groupers.append(grouper)
        if use_xyz:
            mlp_spec[0] += 3
        module.mlps.append(pt_utils.SharedMLP(mlp_spec, bn=bn))


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def _multi_scale(self, xyz, new_xyz, features):
        outs = [_max_over_samples(mlp(grouper(xyz, new_xyz, features)))
                for grouper, mlp in zip(self.groupers, self.mlps)]
        return torch.cat(outs, dim=1)

    def forward(self, xyz, features=None):
        """xyz (B,N,3), features (B,C,N) -> (new_xyz (B,npoint,3), new_features (B,sum C_k,npoint))."""
        new_xyz, _ = _sample_centres(xyz, self.npoint)
        return new_xyz, self._multi_scale(xyz, new_xyz, features)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (op-level path)."""

    def __init__(self, *, npoint, radii, nsamples, mlps, bn=True, use_xyz=True, sample_uniformly=False):
        super().__init__()
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (op-level path)."""

    def __init__(self, *, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetSAModuleVotes(nn.Module):
    """Set abstraction that also returns the sampled indices (VoteNet flavour).

    forward(xyz (B,N,3), features (B,C,N) | None, inds (B,npoint) int32 | None)
        -> (new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds (B,npoint) int32[, unique_cnt])
    """

    def __init__(self, *, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True, pooling='max',
                 sigma=None, normalize_xyz=False, sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.pooling = pooling
        self.mlp_module = None
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else self.radius / 2
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize_xyz,
                sample_uniformly=sample_uniformly, ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp  # same object on purpose: the reference mutates the caller's list too
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)

    def forward(self, xyz, features=None, inds=None):
        if inds is not None:
            assert inds.shape[1] == self.npoint
        if fused.sa_supported(self, xyz, features):
            return fused.sa_forward(self, xyz, features, inds)

        new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        grouped = self.grouper(xyz, new_xyz, features)
        unique_cnt = grouped[2] if self.ret_unique_cnt else None
        grouped_features, grouped_xyz = grouped[0], grouped[1]
        new_features = self.mlp_module(grouped_features)  # (B, mlp[-1], npoint, nsample)
        if self.pooling == 'max':
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'avg':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'rbf':
            # radial-basis weights from the (normalised) local coordinates, averaged over nsample
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)
        if self.ret_unique_cnt:
            return new_xyz, new_features, inds, unique_cnt
        return new_xyz, new_features, inds


class PointnetSAModuleMSGVotes(_PointnetSAModuleBase):
    """Multi-scale set abstraction returning the sampled indices (op-level path)."""

    def __init__(self, *, mlps, npoint, radii, nsamples, bn=True, use_xyz=True, sample_uniformly=False):
        super().__init__()
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)

    def forward(self, xyz, features=None, inds=None):
        new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        return new_xyz, self._multi_scale(xyz, new_xyz, features), inds


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance interpolation from `known` to `unknown` + shared MLP.

    forward(unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n) | None, known_feats (B,C2,m))
        -> (B, mlp[-1], n)
    """

    def __init__(self, *, mlp, bn=True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if fused.fp_supported(self, unknown, known, unknow_feats, known_feats):
            return fused.fp_forward(self, unknown, known, unknow_feats, known_feats)

        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)


class PointnetLFPModuleMSG(nn.Module):
    """Learnable feature propagation from (xyz1, features1) to the points xyz2 (op-level path)."""

    def __init__(self, *, mlps, radii, nsamples, post_mlp, bn=True, use_xyz=True, sample_uniformly=False):
        super().__init__()
        self.post_mlp = pt_utils.SharedMLP(post_mlp, bn=bn)
        _build_scales(self, 0, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)
        del self.npoint  # the reference LFP module has no npoint attribute

    def forward(self, xyz2, xyz1, features2, features1):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            new_features = _max_over_samples(mlp(grouper(xyz1, xyz2, features1)))  # (B, mlp[-1], N2)
            if features2 is not None:
                new_features = torch.cat([new_features, features2], dim=1)
            outs.append(self.post_mlp(new_features.unsqueeze(-1)))
        return torch.cat(outs, dim=1).squeeze(-1)
