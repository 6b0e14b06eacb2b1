"""CPU oracle for the PointNet++ set-abstraction path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product (omni-pq_b200/) never does and fails loudly without its CUDA library.

Two layers:

1. `ext` -- an object with the nine functions of the reference's pybind module `pointnet2._ext`
   (bindings.cpp:11-24; same names, argument order and zero/1e10 initialisation as the .cpp wrappers)
   implemented for CPU tensors by the C restatement in pn2_oracle.c.
2. A restatement of the Python glue above it: the six autograd Functions
   (pointnet2_utils.py:51-291), QueryAndGroup (:294-376), PointnetSAModuleVotes.forward
   (pointnet2_modules.py:210-272), PointnetFPModule.forward (:371-416), SharedMLP
   (pytorch_utils.py:11-36) and the backbone wiring (models/backbone_module.py:33-139), with the
   same parameter names so state_dicts are interchangeable with the product modules.
   The shared-MLP arithmetic is torch CPU fp32 (conv2d / batch_norm / max_pool2d), which is what the
   reference calls at pytorch_utils.py:88-95,43 and pointnet2_modules.py:255.

Parity status: index paths are pinned against the reference's own CUDA kernels (oracle/_ref, run on
the GPU box by tests/test_gpu_ops.py) and the fixtures in tests/golden; the Python glue restated here is
pinned against the reference's own pointnet2/*.py by tests/test_ref_glue_cpu.py; the MLP arithmetic is
"parity unpinned" in the reference (it has no test at that boundary) and is anchored on
torch fp32.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpn2_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pn2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpn2_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.pn2o_opt_n_threads.restype = ctypes.c_int
        _lib.pn2o_num_threads.restype = ctypes.c_int
    return _lib


def _fp(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _ip(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


class _Ext:
    """CPU stand-in for `pointnet2._ext` (bindings.cpp:11-24)."""

    @staticmethod
    def opt_n_threads(n):
        return lib().pn2o_opt_n_threads(int(n))

    @staticmethod
    def furthest_point_sampling(points, nsamples):  # sampling.cpp:72-93
        b, n, _ = points.shape
        out = torch.zeros(b, nsamples, dtype=torch.int32)
        tmp = torch.full((b, n), 1e10, dtype=torch.float32)
        lib().pn2o_furthest_point_sampling(b, n, int(nsamples), _fp(points), _fp(tmp), _ip(out))
        return out

    @staticmethod
    def gather_points(points, idx):  # sampling.cpp:22-46
        b, c, n = points.shape
        m = idx.shape[1]
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().pn2o_gather_points(b, c, n, m, _fp(points), _ip(idx), _fp(out))
        return out

    @staticmethod
    def gather_points_grad(grad_out, idx, n):  # sampling.cpp:48-71
        b, c, m = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().pn2o_gather_points_grad(b, c, int(n), m, _fp(grad_out), _ip(idx), _fp(out))
        return out

    @staticmethod
    def ball_query(new_xyz, xyz, radius, nsample):  # ball_query.cpp:16-40
        b, m, _ = new_xyz.shape
        n = xyz.shape[1]
        idx = torch.zeros(b, m, nsample, dtype=torch.int32)
        lib().pn2o_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _fp(new_xyz), _fp(xyz),
                              _ip(idx))
        return idx

    @staticmethod
    def group_points(points, idx):  # group_points.cpp:19-42
        b, c, n = points.shape
        _, npoints, nsample = idx.shape
        out = torch.zeros(b, c, npoints, nsample, dtype=torch.float32)
        lib().pn2o_group_points(b, c, n, npoints, nsample, _fp(points), _ip(idx), _fp(out))
        return out

    @staticmethod
    def group_points_grad(grad_out, idx, n):  # group_points.cpp:44-67
        b, c, npoints, nsample = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().pn2o_group_points_grad(b, c, int(n), npoints, nsample, _fp(grad_out), _ip(idx), _fp(out))
        return out

    @staticmethod
    def three_nn(unknowns, knows):  # interpolate.cpp:22-48
        b, n, _ = unknowns.shape
        m = knows.shape[1]
        idx = torch.zeros(b, n, 3, dtype=torch.int32)
        dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
        lib().pn2o_three_nn(b, n, m, _fp(unknowns), _fp(knows), _fp(dist2), _ip(idx))
        return dist2, idx

    @staticmethod
    def three_interpolate(points, idx, weight):  # interpolate.cpp:50-78
        b, c, m = points.shape
        n = idx.shape[1]
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().pn2o_three_interpolate(b, c, m, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
        return out

    @staticmethod
    def three_interpolate_grad(grad_out, idx, weight, m):  # interpolate.cpp:79-107
        b, c, n = grad_out.shape
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().pn2o_three_interpolate_grad(b, c, n, int(m), _fp(grad_out), _ip(idx), _fp(weight), _fp(out))
        return out


ext = _Ext()


# ---- autograd glue (pointnet2_utils.py:51-291) -------------------------------------------------
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.shape[2])
        return ext.gather_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        idx, n = ctx.saved
        return ext.gather_points_grad(g.contiguous(), idx, n), None


class _Group(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.shape[2])
        return ext.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, g):
        idx, n = ctx.saved
        return ext.group_points_grad(g.contiguous(), idx, n), None


class _Interp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.saved = (idx, weight, features.shape[2])
        return ext.three_interpolate(features.contiguous(), idx, weight.contiguous())

    @staticmethod
    def backward(ctx, g):
        idx, weight, m = ctx.saved
        return ext.three_interpolate_grad(g.contiguous(), idx, weight, m), None, None


def furthest_point_sample(xyz, npoint):
    return ext.furthest_point_sampling(xyz.detach().contiguous(), npoint)


def gather_operation(features, idx):
    return _Gather.apply(features, idx)


def grouping_operation(features, idx):
    return _Group.apply(features, idx)


def ball_query(radius, nsample, xyz, new_xyz):  # python arg order, pointnet2_utils.py:262,282
    return ext.ball_query(new_xyz.detach().contiguous(), xyz.detach().contiguous(), radius, nsample)


def three_nn(unknown, known):  # pointnet2_utils.py:140-142 returns sqrt(dist2)
    d2, idx = ext.three_nn(unknown.detach().contiguous(), known.detach().contiguous())
    return torch.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    return _Interp.apply(features, idx, weight)


def query_and_group(radius, nsample, xyz, new_xyz, features, use_xyz=True, normalize_xyz=False):
    """QueryAndGroup.forward (pointnet2_utils.py:317-376) without the unused sample_uniformly branch.
    Returns (new_features (B,3+C,m,ns), grouped_xyz (B,3,m,ns), idx)."""
    idx = ball_query(radius, nsample, xyz, new_xyz)
    g_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
    g_xyz = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        g_xyz = g_xyz / radius
    if features is not None:
        g_feat = grouping_operation(features, idx)
        new_features = torch.cat([g_xyz, g_feat], dim=1) if use_xyz else g_feat
    else:
        new_features = g_xyz
    return new_features, g_xyz, idx


# ---- module restatement (same parameter names as the reference => same state_dict keys) ---------
class _BN(nn.Sequential):  # pytorch_utils.py:39-58
    def __init__(self, c):
        super().__init__()
        self.add_module("bn", nn.BatchNorm2d(c))


class _ConvBNReLU(nn.Sequential):  # pytorch_utils.py:67-120,157-188 with bn=True, preact=False
    def __init__(self, cin, cout, bn=True):
        super().__init__()
        conv = nn.Conv2d(cin, cout, kernel_size=(1, 1), bias=not bn)
        nn.init.kaiming_normal_(conv.weight)
        if not bn:
            nn.init.constant_(conv.bias, 0)
        self.add_module("conv", conv)
        if bn:
            self.add_module("bn", _BN(cout))
        self.add_module("activation", nn.ReLU(inplace=True))


class OracleSharedMLP(nn.Sequential):  # pytorch_utils.py:11-36
    def __init__(self, spec, bn=True):
        super().__init__()
        for i in range(len(spec) - 1):
            self.add_module(f"layer{i}", _ConvBNReLU(spec[i], spec[i + 1], bn=bn))


class OracleSAModuleVotes(nn.Module):  # pointnet2_modules.py:164-272 (pooling='max')
    def __init__(self, *, mlp, npoint, radius, nsample, bn=True, use_xyz=True, normalize_xyz=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.use_xyz, self.normalize_xyz = use_xyz, normalize_xyz
        spec = list(mlp)
        if use_xyz and len(spec) > 0:
            spec[0] += 3
        self.mlp_module = OracleSharedMLP(spec, bn=bn)

    def forward(self, xyz, features=None, inds=None):
        if inds is None:
            inds = furthest_point_sample(xyz, self.npoint)
        new_xyz = gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        grouped, _, idx = query_and_group(self.radius, self.nsample, xyz, new_xyz, features,
                                          use_xyz=self.use_xyz, normalize_xyz=self.normalize_xyz)
        self.last_idx = idx
        h = self.mlp_module(grouped)
        h = F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1)
        return new_xyz, h, inds


def fp_weights(dist):  # pointnet2_modules.py:395-397
    r = 1.0 / (dist + 1e-8)
    return r / torch.sum(r, dim=2, keepdim=True)


class OracleFPModule(nn.Module):  # pointnet2_modules.py:356-416
    def __init__(self, *, mlp, bn=True):
        super().__init__()
        self.mlp = OracleSharedMLP(list(mlp), bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        dist, idx = three_nn(unknown, known)
        interp = three_interpolate(known_feats, idx, fp_weights(dist))
        x = torch.cat([interp, unknow_feats], dim=1) if unknow_feats is not None else interp
        return self.mlp(x.unsqueeze(-1)).squeeze(-1)


class OracleBackbone(nn.Module):  # models/backbone_module.py:33-139 (width=2, depth=2)
    def __init__(self, input_feature_dim=0, sa_cls=OracleSAModuleVotes, fp_cls=OracleFPModule):
        super().__init__()
        c0 = input_feature_dim
        self.sa1 = sa_cls(npoint=2048, radius=0.2, nsample=64, mlp=[c0, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa2 = sa_cls(npoint=1024, radius=0.4, nsample=32, mlp=[256, 256, 256, 512], use_xyz=True, normalize_xyz=True)
        self.sa3 = sa_cls(npoint=512, radius=0.8, nsample=16, mlp=[512, 256, 256, 512], use_xyz=True, normalize_xyz=True)
        self.sa4 = sa_cls(npoint=256, radius=1.2, nsample=16, mlp=[512, 256, 256, 512], use_xyz=True, normalize_xyz=True)
        self.fp1 = fp_cls(mlp=[1024, 512, 512])
        self.fp2 = fp_cls(mlp=[1024, 512, 288])

    def forward(self, pointcloud, end_points=None):
        ep = end_points if end_points else {}
        xyz = pointcloud[..., 0:3].contiguous()
        feats = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
        xyz, feats, inds = self.sa1(xyz, feats)
        ep["sa1_inds"], ep["sa1_xyz"], ep["sa1_features"] = inds, xyz, feats
        xyz, feats, inds = self.sa2(xyz, feats)
        ep["sa2_inds"], ep["sa2_xyz"], ep["sa2_features"] = inds, xyz, feats
        xyz, feats, inds = self.sa3(xyz, feats)
        ep["sa3_xyz"], ep["sa3_features"] = xyz, feats
        xyz, feats, inds = self.sa4(xyz, feats)
        ep["sa4_xyz"], ep["sa4_features"] = xyz, feats
        feats = self.fp1(ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], ep["sa4_features"])
        feats = self.fp2(ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], feats)
        ep["fp2_features"] = feats
        ep["fp2_xyz"] = ep["sa2_xyz"]
        ep["fp2_inds"] = ep["sa1_inds"][:, 0:ep["fp2_xyz"].shape[1]]
        ep["seed_inds"], ep["seed_xyz"], ep["seed_features"] = ep["fp2_inds"], ep["fp2_xyz"], ep["fp2_features"]
        return ep


# ---- synthetic inputs: tools/synth_clouds.py (neutral module; re-exported for the tests' convenience) ----
import sys as _sys

if os.path.dirname(_HERE) not in _sys.path:
    _sys.path.insert(0, os.path.dirname(_HERE))
from tools.synth_clouds import scannet_like_cloud, uniform_cloud  # noqa: E402,F401
