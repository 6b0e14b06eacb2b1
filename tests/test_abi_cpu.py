"""The C-ABI library loads without a GPU and exports exactly what include/pn2_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pn2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pn2_b200.h but not exported"


def test_no_undeclared_exports(built_lib):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (pn2_[a-z0-9_]+)", out)))
    assert exported == _declared()


def test_block_size_rule_matches_reference_formula(built_lib):
    """cuda_utils.h:20-24: clamp(2^floor(log2 n), 1, 512); it fixes the FPS tie-break order."""
    from oracle import pn2_oracle as O
    lib = ctypes.CDLL(built_lib)
    lib.pn2_ref_block_size.restype = ctypes.c_int
    for n in [1, 2, 3, 127, 128, 129, 200, 255, 256, 511, 512, 513, 777, 1024, 40000, 50000, 1 << 20]:
        want = min(512, 1 << (n.bit_length() - 1))
        assert lib.pn2_ref_block_size(n) == want == O.ext.opt_n_threads(n)


def test_argument_errors_are_codes_not_exits(built_lib):
    """Bad extents / null pointers return PN2_ERR_INVALID_ARG with a message; nothing calls exit()
    (the reference's CUDA_CHECK_ERRORS does, cuda_utils.h:35-44).  No kernel is launched here."""
    lib = ctypes.CDLL(built_lib)
    lib.pn2_last_error.restype = ctypes.c_char_p
    assert lib.pn2_version() >= 100
    rc = lib.pn2_furthest_point_sampling(1, 0, 4, None, None, None, None, None)
    assert rc == -1 and b"pn2_furthest_point_sampling" in lib.pn2_last_error()
    rc = lib.pn2_furthest_point_sampling(1, 16, 4, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.pn2_last_error()
    assert lib.pn2_ball_query(1, 16, 4, ctypes.c_float(0.1), 0, None, None, None, None) == -1
    # every nsample is accepted like in the reference (ball_query_gpu.cu has no cap): only the null pointers are refused
    assert lib.pn2_ball_query(1, 16, 4, ctypes.c_float(0.1), 100000, None, None, None, None) == -1
    assert b"null" in lib.pn2_last_error()
    assert lib.pn2_three_nn(-1, 1, 1, None, None, None, None, None) == -1
    assert lib.pn2_group_points(1, 4, 16, 1 << 20, 1 << 20, None, None, None, None) == -1  # int32 overflow
    # empty problems are fine and touch nothing
    assert lib.pn2_furthest_point_sampling(0, 16, 4, None, None, None, None, None) == 0
    assert lib.pn2_gather_points(0, 3, 16, 4, None, None, None, None) == 0
    assert lib.pn2_three_interpolate(2, 0, 4, 4, None, None, None, None, None) == 0


def test_wgrad_workspace_covers_both_tilings(built_lib):
    """pn2_mlp_wgrad_workspace is what a binder allocates BEFORE it knows the row sources, so it must cover the larger of
    the two slice counts: the plain tiling and, for kp = 128 m + 4 (gathered source, xyz block folded into the last
    feature tile -- one n-tile fewer, hence more position slices), the folded one.  Host arithmetic only."""
    lib = ctypes.CDLL(built_lib)
    lib.pn2_mlp_wgrad_workspace.restype = ctypes.c_longlong
    sms = 148  # without a device the library assumes a B200
    for rows, np_, kp in [(131072, 128, 8), (32768, 256, 260), (8192, 256, 516), (4096, 256, 516), (370, 48, 132), (1, 4, 4)]:
        ws = lib.pn2_mlp_wgrad_workspace(rows, np_, kp)
        assert ws > 0 and ws % (np_ * kp) == 0
        slices = ws // (np_ * kp)
        tiles_m, tiles_n = -(-np_ // 128), -(-kp // 128)
        cap = max(1, -(-rows // 128))                    # a slice is at least 128 positions
        plain = max(1, min(sms // (tiles_m * tiles_n), cap, 512))
        assert slices >= plain
        if kp > 128 and kp % 128 == 4:
            folded = max(1, min(sms // (tiles_m * (tiles_n - 1)), cap, 512))
            assert slices >= folded > 0
        assert slices * tiles_m * max(tiles_n - 1, 1) <= max(sms, tiles_m * tiles_n)  # still one wave of CTAs


def test_python_binding_rejects_cpu_tensors(built_lib):
    """Same contract as the reference wrappers: CPU tensors -> RuntimeError ("CPU not supported",
    sampling.cpp:41), wrong dtype / layout -> RuntimeError (utils.h:10-30).  There is no fallback."""
    import torch
    import _pn2
    xyz = torch.rand(1, 32, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _pn2.furthest_point_sampling(xyz, 4)
    with pytest.raises(RuntimeError, match="contiguous"):
        _pn2.gather_points(torch.rand(1, 32, 3).transpose(1, 2), torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="int tensor"):
        _pn2.gather_points(torch.rand(1, 3, 32), torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="float tensor"):
        _pn2.ball_query(torch.rand(1, 4, 3).double(), xyz, 0.1, 4)
    import pointnet2_utils
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pointnet2_utils.furthest_point_sample(xyz, 4)
