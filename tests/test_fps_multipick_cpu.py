"""The multi-pick FPS scheme of fps_multipick_kernel (several samples resolved per exchange round, see the
comment block in omni-pq_b200/csrc/fps.cu) must reproduce the reference's one-pick-per-round order bit for bit.
tests/emul/fps_multipick_emul.c restates the kernel's thread/warp/CTA mapping, candidate + bound construction and
resolution loop in C; here it is checked against the oracle on every parity case (ties, skipped points, ragged
sizes, tiny clouds) and for every cluster size, without a GPU.  The GPU suite checks the kernel itself."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
from oracle import pn2_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "fps_multipick_emul.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "emul", "fps_multipick_emul.c"), "-lm"])
    lib = ctypes.CDLL(so)
    lib.fps_multipick_emul.restype = ctypes.c_int
    lib.fps_multipick_emul.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]

    lib.fps_prefix_verify_emul.restype = ctypes.c_int
    lib.fps_prefix_verify_emul.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p]

    def verify(xyz, m):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = xyz.shape[0]
        bs = max(1, min(512, 2 ** int(math.log(n) / math.log(2.0))))
        return lib.fps_prefix_verify_emul(n, m, bs, xyz.ctypes.data)

    def run(xyz, m, cs):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = xyz.shape[0]
        bs = max(1, min(512, 2 ** int(math.log(n) / math.log(2.0))))  # cuda_utils.h:20-24
        idx = np.zeros(m, dtype=np.int32)
        rounds = lib.fps_multipick_emul(n, m, cs, bs, xyz.ctypes.data, idx.ctypes.data)
        assert rounds >= 0
        return idx, rounds
    run.verify = verify
    return run


@pytest.mark.parametrize("name", cases.SMALL)
@pytest.mark.parametrize("cs", [1, 2, 16])
def test_multipick_equals_reference_order_small_cases(name, cs, emul):
    make, stages = cases.CASES[name]
    xyz = make()
    m = [s for s in stages if s[0] == "fps"][0][1]
    want = O.furthest_point_sample(xyz, m).numpy()
    for b in range(xyz.shape[0]):
        got, _ = emul(xyz[b].numpy(), m, cs)
        assert np.array_equal(got, want[b]), (name, cs, b, int(np.argmax(got != want[b])))


def test_multipick_more_samples_than_distinct_points(emul):
    """m close to n with 25 % duplicates: min-distances reach exactly 0 and the arg-max is decided by the
    tie-break rank alone."""
    xyz = cases.c1_duplicates()[:1, :300].contiguous()
    xyz[0, 200:] = xyz[0, :100]
    want = O.furthest_point_sample(xyz, 290).numpy()[0]
    for cs in (1, 4):
        got, _ = emul(xyz[0].numpy(), 290, cs)
        assert np.array_equal(got, want), cs


def test_multipick_all_points_skipped(emul):
    xyz = torch.full((1, 64, 3), 0.001)
    want = O.furthest_point_sample(xyz, 16).numpy()[0]
    got, _ = emul(xyz[0].numpy(), 16, 1)
    assert np.array_equal(got, want) and not got.any()


def test_multipick_scannet_levels_and_round_count(emul):
    """configs[1]: the 40k-point cloud and the three coarser levels; the scheme needs ~9x fewer exchange rounds
    than samples at the first level."""
    cloud = cases.c2_scannet()
    xyz = cloud
    for m, cs in ((2048, 16), (1024, 1), (512, 1), (256, 1)):
        want = O.furthest_point_sample(xyz, m)
        got, rounds = emul(xyz[0].numpy(), m, cs)
        assert np.array_equal(got, want.numpy()[0]), m
        if m == 2048:
            assert rounds < 2047 // 5, rounds
        xyz = torch.gather(xyz, 1, want.long()[..., None].expand(-1, -1, 3)).contiguous()


def _fps_ordered(xyz, m):
    inds = O.furthest_point_sample(xyz, m)
    return torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()


@pytest.mark.parametrize("name", cases.SMALL + ["c2_scannet"])
def test_prefix_order_check_passes_only_when_the_answer_is_the_identity(name, emul):
    """The prefix-order shortcut (fps.cu) answers 0..m-1 when its parallel check passes.  On unordered clouds and on
    FPS-ordered ones (what every level after the first samples from) -- including duplicates, lattice ties and
    skipped points, where the subset's tie-break ranks can break the identity -- the check must pass exactly when
    the oracle's samples ARE 0..m-1."""
    make, stages = cases.CASES[name]
    xyz = make()
    m0 = [s for s in stages if s[0] == "fps"][0][1]
    passed = 0
    for cloud, m in ((xyz, m0), (_fps_ordered(xyz, m0), m0 // 2), (_fps_ordered(xyz, m0), m0 - 1)):
        want = O.furthest_point_sample(cloud, m).numpy()
        for b in range(cloud.shape[0]):
            identity = np.array_equal(want[b], np.arange(m, dtype=want.dtype))
            ok = emul.verify(cloud[b].numpy(), m) == 0
            assert ok == identity, (name, b, m, ok, identity)
            passed += ok
    if name in ("c1_uniform", "c2_scannet"):
        assert passed >= 2  # the ordered levels of an ordinary cloud take the shortcut


@pytest.mark.parametrize("seed", range(24))
def test_multipick_and_prefix_check_on_random_tie_heavy_clouds(seed, emul):
    """Randomised adversarial inputs: points on a coarse integer lattice (massive exact distance ties), random
    duplicates, a few points inside the skip radius, ragged sizes that change the reference block size -- the
    multi-pick scheme must reproduce the oracle's order for every cluster size, and the prefix-order check must
    pass exactly when the oracle's answer on the FPS-ordered cloud is the identity."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    n = int(rng.integers(40, 700))
    m = int(rng.integers(2, n))
    pts = rng.integers(0, 6, size=(n, 3)).astype(np.float32) * np.float32(0.25) + np.float32(0.125)
    dup = rng.random(n) < 0.2
    pts[dup] = pts[rng.integers(0, n, size=int(dup.sum()))]
    near = rng.random(n) < 0.03
    pts[near] = (rng.random((int(near.sum()), 3)) * 0.02).astype(np.float32)
    xyz = torch.from_numpy(pts)[None].contiguous()
    want = O.furthest_point_sample(xyz, m).numpy()[0]
    for cs in (1, 2, 8):
        got, _ = emul(pts, m, cs)
        assert np.array_equal(got, want), (seed, n, m, cs, int(np.argmax(got != want)))
    ordered = torch.gather(xyz, 1, torch.from_numpy(want).long()[None, :, None].expand(-1, -1, 3)).contiguous()
    for cloud, mm in ((xyz, m), (ordered, max(2, m // 2)), (ordered, m - 1 if m > 2 else 2)):
        if mm >= cloud.shape[1]:
            continue
        ref = O.furthest_point_sample(cloud, mm).numpy()[0]
        identity = np.array_equal(ref, np.arange(mm, dtype=ref.dtype))
        assert (emul.verify(cloud[0].numpy(), mm) == 0) == identity, (seed, n, mm, identity)
