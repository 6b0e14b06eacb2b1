"""Stage the reference's *caller* files under baseline/_ref/ (git-ignored, travels to the GPU box).

SURVEY.md 8(c): /root/reference does not exist on the B200 box, so everything a GPU parity test needs
from the reference has to ride along in the repository snapshot.  This script copies, verbatim and
without editing a byte, the files that *call* the hot path

    pointnet2/*.py                 (the reference's own Python glue: pointnet2_modules / _utils / pytorch_utils)
    models/  utils/                (Pointnet2Backbone, PQ_Transformer, FPSModule, VotingModule, transformer ...)
    scannet/model_util_scannet.py  scannet/meta_data/scannet_means.npz

into baseline/_ref/.  The directory is listed in .gitignore (never part of the history, never product
source) and NOT in .gpurunignore.  Consumers: tests/ (real callers on our modules vs on the reference's
own kernels, oracle/_ref/pn2_ref_ext.so) and bench.py (imports the real models/backbone_module.py when it
is staged).  The compiled reference kernels themselves come from oracle/build_ref.py.

Usage:  python tools/stage_reference.py [--force]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

_DIRS = ("models", "utils")
_FILES = ("scannet/model_util_scannet.py", "scannet/meta_data/scannet_means.npz")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "pointnet2"))


def staged() -> bool:
    return os.path.exists(os.path.join(DST, "models", "pq_transformer.py"))


def stage(force: bool = False) -> str:
    if not available():
        if staged():
            return DST
        raise FileNotFoundError(f"{REF} not present and nothing staged under {DST}")
    if staged() and not force:
        return DST
    os.makedirs(DST, exist_ok=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for d in _DIRS:
        shutil.copytree(os.path.join(REF, d), os.path.join(DST, d), dirs_exist_ok=True, ignore=ignore)
    pn2 = os.path.join(DST, "pointnet2")
    os.makedirs(pn2, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "pointnet2")):
        if f.endswith(".py") and f != "setup.py":
            shutil.copy2(os.path.join(REF, "pointnet2", f), os.path.join(pn2, f))
    for f in _FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copy2(os.path.join(REF, f), os.path.join(DST, f))
    return DST


def root():
    """Directory holding the reference's callers: /root/reference here, baseline/_ref on the GPU box; None if neither."""
    if available():
        return REF
    return DST if staged() else None


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
