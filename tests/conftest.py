import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "omni-pq_b200")
for p in (ROOT, os.path.join(ROOT, "tests"), PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libpn2_b200.so, building it if nvcc is available and it is missing/stale."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pn2_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(mod.NVCC):
        return mod.build()
    assert os.path.exists(mod.SO), "libpn2_b200.so missing and no nvcc to build it"
    return mod.SO


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    out = {}
    for kind in ("oracle", "ref"):
        path = os.path.join(ROOT, "tests", "golden", f"pn2_golden_{kind}.npz")
        if os.path.exists(path):
            out[kind] = dict(np.load(path))
    return out
