"""The float-path oracle is "the reference's own glue over the pinned CPU ops": tests/ref_glue_check.py imports the
reference's unchanged pointnet2_modules.py / pointnet2_utils.py / pytorch_utils.py and models/backbone_module.py
(from /root/reference, or its staged copy baseline/_ref) with the CPU oracle's ext installed as `pointnet2._ext`
and asserts torch.equal against oracle.pn2_oracle's restated modules: forward, indices, every gradient (incl. the
xyz gradient of vote_aggregation's SA) and the BatchNorm buffers."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
from tools import stage_reference  # noqa: E402


@pytest.mark.skipif(stage_reference.root() is None, reason="reference glue neither at /root/reference nor staged")
def test_reference_glue_over_oracle_ext_equals_oracle_modules_bitwise():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_glue_check.py")], capture_output=True,
                       text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "bit-exact checks passed" in r.stdout


def test_stage_script_lists_only_callers():
    """The staging recipe copies the callers + the reference's Python glue, never the kernels' sources."""
    src = open(os.path.join(ROOT, "tools", "stage_reference.py")).read()
    assert "_ext_src" not in src
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "baseline/_ref/" in gi
    ign = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(ign) or "baseline" not in open(ign).read()
