"""Per-op device timings (CUDA events) of libpn2_b200.so next to the reference's own kernels
(oracle/_ref/pn2_ref_ext.so) on the backbone's shapes.  Development aid; bench.py is the contract.

    python tools/op_bench.py [--batch 1] [--out gpurun_out/op_bench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "omni-pq_b200"))

from oracle import pn2_oracle as O  # noqa: E402  (input generator only)
from oracle import build_ref  # noqa: E402
import _pn2  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    B = args.batch
    ref = build_ref.load()
    clouds = torch.stack([O.scannet_like_cloud(40000, seed=1234 + i)[:, :3] for i in range(B)]).contiguous().cuda()
    rows = []

    def both(name, ours, theirs):
        t_o = timeit(ours)
        t_r = timeit(theirs, iters=5, warm=1) if (ref is not None and theirs is not None) else None
        rows.append({"op": name, "ours_us": round(t_o, 2), "ref_us": None if t_r is None else round(t_r, 2),
                     "speedup": None if t_r is None else round(t_r / t_o, 2)})
        print(rows[-1], flush=True)

    xyz = clouds
    for lvl, (npoint, radius, ns) in enumerate([(2048, 0.2, 64), (1024, 0.4, 32), (512, 0.8, 16), (256, 1.2, 16)], 1):
        n = xyz.shape[1]
        both(f"fps_{n}_to_{npoint}", lambda: _pn2.furthest_point_sampling(xyz, npoint),
             (lambda: ref.furthest_point_sampling(xyz, npoint)) if ref else None)
        inds, new_xyz = _pn2.furthest_point_sampling(xyz, npoint, return_xyz=True)
        both(f"ball_query_{n}x{npoint}_ns{ns}", lambda: _pn2.ball_query(new_xyz, xyz, radius, ns),
             (lambda: ref.ball_query(new_xyz, xyz, radius, ns)) if ref else None)
        idx = _pn2.ball_query(new_xyz, xyz, radius, ns)
        c = [3, 256, 512, 512][lvl - 1]
        feats = torch.randn(B, c, n, device="cuda")
        both(f"group_points_c{c}_{npoint}x{ns}", lambda: _pn2.group_points(feats, idx),
             (lambda: ref.group_points(feats, idx)) if ref else None)
        g = torch.randn(B, c, npoint, ns, device="cuda")
        both(f"group_points_grad_c{c}_{npoint}x{ns}", lambda: _pn2.group_points_grad(g, idx, n),
             (lambda: ref.group_points_grad(g, idx, n)) if ref else None)
        xyz = new_xyz
    for (n, m, c) in [(512, 256, 512), (1024, 512, 512), (50000, 2048, 256)]:
        u, k = torch.rand(B, n, 3, device="cuda"), torch.rand(B, m, 3, device="cuda")
        both(f"three_nn_{n}x{m}", lambda: _pn2.three_nn(u, k), (lambda: ref.three_nn(u, k)) if ref else None)
        d2, idx = _pn2.three_nn(u, k)
        w = torch.rand(B, n, 3, device="cuda")
        f = torch.randn(B, c, m, device="cuda")
        both(f"three_interpolate_c{c}_{n}", lambda: _pn2.three_interpolate(f, idx, w),
             (lambda: ref.three_interpolate(f, idx, w)) if ref else None)
        g = torch.randn(B, c, n, device="cuda")
        both(f"three_interpolate_grad_c{c}_{n}", lambda: _pn2.three_interpolate_grad(g, idx, w, m),
             (lambda: ref.three_interpolate_grad(g, idx, w, m)) if ref else None)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"batch": B, "gpu": torch.cuda.get_device_name(0), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
