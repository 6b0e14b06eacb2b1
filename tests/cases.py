"""Seeded synthetic inputs shared by the golden generator and the parity tests.

Every case is a dict name -> callable returning `(xyz (B,N,3) float32 tensor, stages)` where `stages`
lists the index ops run on it: ("fps", m), ("ball", radius, nsample) on the previous stage's centres,
("three_nn",) from the cloud to the current centres.  Inputs come from numpy PCG64 so that the CPU
oracle, the reference extension and the B200 kernels all see identical bits.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pn2_oracle as O  # noqa: E402


def c1_uniform():
    xyz, _ = O.uniform_cloud(2, 1024, 3, seed=0)
    return xyz


def c1_duplicates():
    """25 % of the points are exact copies of other points: FPS / 3-NN ties (tree tie-break order)."""
    xyz, _ = O.uniform_cloud(2, 1024, 3, seed=1)
    xyz = xyz.clone()
    xyz[:, 768:] = xyz[:, :256]
    return xyz


def c1_near_origin():
    """3 % of the points lie within sqrt(1e-3) of the origin: FPS must skip them (sampling_gpu.cu:105-106),
    including index 0 itself in cloud 1."""
    xyz, _ = O.uniform_cloud(2, 1024, 3, seed=2)
    xyz = xyz.clone() - 0.5
    rng = np.random.Generator(np.random.PCG64(22))
    tiny = torch.from_numpy((rng.random((2, 32, 3)) * 0.03 - 0.015).astype(np.float32))
    xyz[:, 100:132] = tiny
    xyz[1, 0] = torch.tensor([0.001, -0.002, 0.0005])
    return xyz


def c1_grid():
    """Points on a regular lattice: massive exact distance ties in FPS, ball query and 3-NN."""
    g = torch.arange(10, dtype=torch.float32) * 0.125 + 0.25
    pts = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), dim=-1).reshape(-1, 3)  # 1000 points
    perm = torch.from_numpy(np.random.Generator(np.random.PCG64(3)).permutation(1000))
    return torch.stack([pts, pts[perm]]).contiguous()


def c1_ragged():
    """N not a multiple of anything convenient (block size 512 with a ragged last stripe)."""
    xyz, _ = O.uniform_cloud(3, 777, 3, seed=4)
    return xyz


def c1_small():
    """N = 200 -> the reference launches 128 threads: a different tie-break tree."""
    xyz, _ = O.uniform_cloud(2, 200, 3, seed=5)
    xyz = xyz.clone()
    xyz[:, 150:] = xyz[:, :50]
    return xyz


def c2_scannet():
    return O.scannet_like_cloud(40000, seed=1234)[None, :, :3].contiguous()


def c5_arkit():
    return O.scannet_like_cloud(50000, seed=4321, centred=True, yaw=True)[None, :, :3].contiguous()


CASES = {
    "c1_uniform": (c1_uniform, [("fps", 512), ("ball", 0.2, 64), ("three_nn",)]),
    "c1_duplicates": (c1_duplicates, [("fps", 512), ("ball", 0.2, 64), ("three_nn",)]),
    "c1_near_origin": (c1_near_origin, [("fps", 256), ("ball", 0.1, 16), ("three_nn",)]),
    "c1_grid": (c1_grid, [("fps", 300), ("ball", 0.13, 8), ("ball", 0.3, 32), ("three_nn",)]),
    "c1_ragged": (c1_ragged, [("fps", 100), ("ball", 0.25, 20), ("three_nn",)]),
    "c1_small": (c1_small, [("fps", 64), ("ball", 0.3, 16), ("three_nn",)]),
    "c2_scannet": (c2_scannet, [("fps", 2048), ("ball", 0.2, 64), ("three_nn",), ("fps", 1024), ("ball", 0.4, 32),
                                ("fps", 512), ("ball", 0.8, 16), ("fps", 256), ("ball", 1.2, 16)]),
    "c5_arkit": (c5_arkit, [("fps", 2048), ("ball", 0.2, 64), ("three_nn",)]),
}
SMALL = [k for k in CASES if k.startswith("c1_")]


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()[:16]


def run_case(name, ext, device="cpu"):
    """Run the index ops of a case through `ext` (an object with the reference's `_ext` functions:
    the CPU oracle, the compiled reference extension or the ctypes binding of libpn2_b200.so).
    Returns an ordered dict key -> CPU tensor."""
    make, stages = CASES[name]
    xyz0 = make().to(device)
    out = {}
    cloud, centres = xyz0, None
    level = 0
    for st in stages:
        if st[0] == "fps":
            if centres is not None:
                cloud = centres  # next abstraction level samples from the previous level's centres
            level += 1
            inds = ext.furthest_point_sampling(cloud, st[1])
            flipped = cloud.transpose(1, 2).contiguous()
            centres = ext.gather_points(flipped, inds).transpose(1, 2).contiguous()
            out[f"l{level}_fps{st[1]}"] = inds
        elif st[0] == "ball":
            out[f"l{level}_ball_r{st[1]}_ns{st[2]}"] = ext.ball_query(centres, cloud, st[1], st[2])
        elif st[0] == "three_nn":
            d2, idx = ext.three_nn(cloud, centres)
            out[f"l{level}_nn_idx"] = idx
            out[f"l{level}_nn_dist2"] = d2
    return {k: v.cpu() for k, v in out.items()}, digest(xyz0)
