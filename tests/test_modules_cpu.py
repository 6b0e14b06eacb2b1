"""Host-side mirror of the reference module interface: names, constructor behaviour, state_dict keys."""
import json
import os
import sys

import pytest
import torch

from conftest import PKG, ROOT


def test_backbone_state_dict_keys_match_reference(built_lib):
    """Keys/shapes dumped from the reference's own Pointnet2Backbone(input_feature_dim=3)
    (tests/golden/backbone_state_dict_keys.json) -- checkpoints must load unchanged."""
    from backbone import Pointnet2Backbone
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "backbone_state_dict_keys.json")))
    got = {k: list(v.shape) for k, v in Pointnet2Backbone(input_feature_dim=3).state_dict().items()}
    assert list(got) == list(want)
    assert all(tuple(got[k]) == tuple(want[k]) for k in want)


def test_constructor_mutates_caller_mlp_like_reference(built_lib):
    """pointnet2_modules.py:204-206: `mlp[0] += 3` happens on the caller's list."""
    import pointnet2_modules as M
    spec = [3, 64]
    sa = M.PointnetSAModuleVotes(npoint=512, radius=0.2, nsample=64, mlp=spec, use_xyz=True, normalize_xyz=True)
    assert spec == [6, 64]
    assert sa.mlp_module.layer0.conv.weight.shape == (64, 6, 1, 1)
    assert sa.mlp_module.layer0.conv.bias is None
    assert sa.sigma == 0.1 and sa.pooling == "max"
    assert isinstance(sa.mlp_module.layer0.bn.bn, torch.nn.BatchNorm2d)
    fp = M.PointnetFPModule(mlp=[8, 4])
    assert list(fp.state_dict())[0] == "mlp.layer0.conv.weight"


def test_all_reference_classes_present(built_lib):
    import pointnet2_modules as M
    import pointnet2_utils as U
    import pytorch_utils as P
    for n in ["_PointnetSAModuleBase", "PointnetSAModuleMSG", "PointnetSAModule", "PointnetSAModuleVotes",
              "PointnetSAModuleMSGVotes", "PointnetFPModule", "PointnetLFPModuleMSG"]:
        assert hasattr(M, n)
    for n in ["RandomDropout", "FurthestPointSampling", "furthest_point_sample", "GatherOperation", "gather_operation",
              "ThreeNN", "three_nn", "ThreeInterpolate", "three_interpolate", "GroupingOperation",
              "grouping_operation", "BallQuery", "ball_query", "QueryAndGroup", "GroupAll"]:
        assert hasattr(U, n)
    for n in ["SharedMLP", "BatchNorm1d", "BatchNorm2d", "BatchNorm3d", "Conv1d", "Conv2d", "Conv3d", "FC",
              "set_bn_momentum_default", "BNMomentumScheduler"]:
        assert hasattr(P, n)
    msg = M.PointnetSAModuleMSG(npoint=2, radii=[5.0, 10.0], nsamples=[6, 3], mlps=[[9, 3], [9, 6]])
    assert list(msg.state_dict())[0] == "mlps.0.layer0.conv.weight"
    lfp = M.PointnetLFPModuleMSG(mlps=[[4, 8]], radii=[0.5], nsamples=[4], post_mlp=[8, 8])
    assert "post_mlp.layer0.conv.weight" in lfp.state_dict()
    pre = P.SharedMLP([4, 8, 8], bn=True, preact=True, first=True)
    assert list(pre.layer0._modules) == ["conv"] and list(pre.layer1._modules) == ["bn", "activation", "conv"]


def test_same_seed_same_init_as_oracle_modules(built_lib):
    """Parameter creation order and initialisers follow pytorch_utils.py:157-188, so the same seed
    gives the same weights as the oracle's restatement (and as the reference)."""
    from backbone import Pointnet2Backbone
    from oracle import pn2_oracle as O
    torch.manual_seed(0)
    ours = Pointnet2Backbone(input_feature_dim=3).state_dict()
    torch.manual_seed(0)
    theirs = O.OracleBackbone(input_feature_dim=3).state_dict()
    assert list(ours) == list(theirs)
    assert all(torch.equal(ours[k], theirs[k]) for k in ours)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree only exists in the build container")
def test_reference_callers_load_our_modules_unchanged(built_lib):
    """models/backbone_module.py:15-18 appends ROOT/pointnet2 to sys.path and imports
    pointnet2_modules; with our directory earlier on sys.path it must pick up ours."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(1, '/root/reference');\n"
        "from models.backbone_module import Pointnet2Backbone\n"
        "import pointnet2_modules, pointnet2_utils, pytorch_utils\n"
        "assert pointnet2_modules.__file__.startswith(%r), pointnet2_modules.__file__\n"
        "assert pointnet2_utils.__file__.startswith(%r)\n"
        "m = Pointnet2Backbone(input_feature_dim=3)\n"
        "import models.utils.pointnet_util as pu\n"
        "assert pu.pointnet2_utils is pointnet2_utils\n"
        "print(type(m.sa1).__module__, len(m.state_dict()))\n" % (PKG, PKG, PKG))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == ["pointnet2_modules", "96"]
