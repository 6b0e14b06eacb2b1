"""Benchmark/test harness wiring of the PQ-Transformer backbone on top of `pointnet2_modules`.

The reference's `models/backbone_module.py:33-139` is a *caller* of the hot path and runs unchanged on
these modules (tests/test_modules_cpu.py imports it from /root/reference when present).  That file is not
available on the GPU box, so bench.py and the parity tests use this equivalent wiring of the same
constants (SA1 2048/0.2/64, SA2 1024/0.4/32, SA3 512/0.8/16, SA4 256/1.2/16, FP1, FP2; width=2, depth=2),
with identical attribute names and therefore identical state_dict keys and end_points.
"""
import torch.nn as nn

from pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes

SA_SPECS = (  # (npoint, radius, nsample, mlp widths after the input width)
    (2048, 0.2, 64, (128, 128, 256)),
    (1024, 0.4, 32, (256, 256, 512)),
    (512, 0.8, 16, (256, 256, 512)),
    (256, 1.2, 16, (256, 256, 512)),
)


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=0, sa_cls=PointnetSAModuleVotes, fp_cls=PointnetFPModule):
        super().__init__()
        c_in = input_feature_dim
        for i, (npoint, radius, nsample, widths) in enumerate(SA_SPECS, start=1):
            setattr(self, f"sa{i}", sa_cls(npoint=npoint, radius=radius, nsample=nsample, mlp=[c_in, *widths],
                                           use_xyz=True, normalize_xyz=True))
            c_in = widths[-1]
        self.fp1 = fp_cls(mlp=[512 + 512, 512, 512])
        self.fp2 = fp_cls(mlp=[512 + 512, 512, 288])

    def forward(self, pointcloud, end_points=None):
        ep = end_points if end_points else {}
        xyz = pointcloud[..., 0:3].contiguous()
        features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
        for i in range(1, 5):
            xyz, features, inds = getattr(self, f"sa{i}")(xyz, features)
            if i <= 2:
                ep[f"sa{i}_inds"] = inds
            ep[f"sa{i}_xyz"], ep[f"sa{i}_features"] = xyz, features
        features = self.fp1(ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], ep["sa4_features"])
        features = self.fp2(ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], features)
        ep["fp2_features"] = features
        ep["fp2_xyz"] = ep["sa2_xyz"]
        ep["fp2_inds"] = ep["sa1_inds"][:, 0:ep["fp2_xyz"].shape[1]]
        ep["seed_inds"], ep["seed_xyz"], ep["seed_features"] = ep["fp2_inds"], ep["fp2_xyz"], ep["fp2_features"]
        return ep
