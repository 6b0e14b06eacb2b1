"""The reference's REAL caller files on the GPU (BASELINE.json configs[1]/[2]):
models/backbone_module.py and models/pq_transformer.py (with FPSModule, VotingModule, vote_aggregation, the
transformer decoder and the prediction heads), imported unchanged from baseline/_ref (staged by
tools/stage_reference.py; /root/reference in the build container), run twice in fresh processes by
tests/ref_callers.py:

    ref   the reference's own pointnet2/*.py over the reference's own CUDA kernels (oracle/_ref/pn2_ref_ext.so) and
          torch/cuDNN fp32 (allow_tf32=False)
    ours  the same model files with omni-pq_b200/ first on sys.path (our modules on libpn2_b200.so)

with the SAME state_dict (written by the first process, loaded strictly by the second) and the same clouds."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, ROOT)
from tools import stage_reference  # noqa: E402

needs_ref = pytest.mark.skipif(
    stage_reference.root() is None or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pn2_ref_ext.so")),
    reason="reference callers (baseline/_ref) or reference kernels (oracle/_ref) not staged")


def _run(impl, case, out, extra):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_callers.py"), "--impl", impl, "--case", case, "--out", out] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    return dict(np.load(out))


def _pair(case, tmp_path, extra=()):
    state = str(tmp_path / "state.pt")
    ref = _run("ref", case, str(tmp_path / "ref.npz"), ["--save-state", state, *extra])
    ours = _run("ours", case, str(tmp_path / "ours.npz"), ["--load-state", state, *extra])
    assert "_ref/pointnet2" in str(ref["modules_file"]) or "/root/reference/pointnet2" in str(ref["modules_file"])
    assert "omni-pq_b200" in str(ours["modules_file"])
    return ref, ours


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@needs_ref
def test_real_backbone_train_step_matches_reference_kernels(tmp_path):
    """configs[1]: models/backbone_module.Pointnet2Backbone(input_feature_dim=3).train(), one 40000 x 6 cloud, fwd+bwd.
    Index paths bit-exact through all levels; sa1 features 1e-5; deeper features within the chain tolerance (18
    training-mode BatchNorm layers evaluated by two fp32 implementations, each feeding on its own slightly different
    outputs: measured 1.4e-6 at sa1 growing to 3.5e-5 at fp2); BatchNorm running statistics 2e-5; gradients in
    relative L2 (ReLU / max-pool mask flips, arbitrated against fp64 in test_gpu_fp64_arbiter.py)."""
    ref, ours = _pair("backbone_train", tmp_path)
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
        assert np.array_equal(ref[k], ours[k]), k
    assert rel(ours["sa1_features"], ref["sa1_features"]) <= 1e-5
    for k in ("sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        assert rel(ours[k], ref[k]) <= 1e-4, (k, rel(ours[k], ref[k]))
    for k in ref:
        if k.startswith("buf."):
            if ref[k].dtype.kind in "iu":
                assert np.array_equal(ref[k], ours[k]), k
            else:
                assert rel(ours[k], ref[k]) <= 2e-5, (k, rel(ours[k], ref[k]))
        if k.startswith("grad."):
            assert rel_l2(ours[k], ref[k]) <= 3e-2, (k, rel_l2(ours[k], ref[k]))


@needs_ref
def test_pq_transformer_end_to_end_batch8_quad_logits(tmp_path):
    """configs[2]: PQ_Transformer(num_class=18, num_heading_bin=1, num_size_cluster=18, scannet means, 256/256
    proposals, 'vote').eval() on 8 x 40000-point clouds: quad logits (`last_quad_scores`) within 1e-4 of the reference.

    The layout (quad) branch samples the seeds (= sa2_xyz, bit-exact), so all of its outputs must agree.  The object
    branch runs FPS over `vote_xyz`, a FLOAT output of VotingModule (agrees to ~1e-6): a near-tie between two
    candidates can resolve differently for two fp32 implementations (it does for the reference itself across GPUs /
    cuDNN versions), after which that cloud's proposals are a different -- equally valid -- sample (measured: 1 of 8
    clouds, two picks swapped).  Hence: object outputs are compared on the clouds whose vote-FPS picks coincide, and
    those must be the large majority."""
    ref, ours = _pair("pq_eval", tmp_path, ["--batch", "8"])
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "aggregated_sample_xyz"):
        assert np.array_equal(ref[k], ours[k]), k
    assert rel(ours["fp2_features"], ref["fp2_features"]) <= 1e-4
    assert np.abs(ours["last_quad_scores"] - ref["last_quad_scores"]).max() <= 1e-4       # the configs[2] criterion
    assert rel(ours["last_quad_center"], ref["last_quad_center"]) <= 1e-5
    assert np.abs(ours["vote_xyz"] - ref["vote_xyz"]).max() <= 1e-4                      # metres
    same = [b for b in range(8) if np.abs(ours["aggregated_vote_xyz"][b] - ref["aggregated_vote_xyz"][b]).max() <= 1e-4]
    assert len(same) >= 6, f"vote-FPS picks coincide on only {len(same)} of 8 clouds"
    for k in ("last_objectness_scores", "last_sem_cls_scores", "cluster_feature", "last_center"):
        d = np.abs(ours[k][same] - ref[k][same]).max()
        assert d <= 1e-4, (k, d)
