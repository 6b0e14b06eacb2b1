"""Micro-benchmark of the shared-MLP GEMM kernels on the backbone's shapes (development aid).
    python tools/gemm_bench.py [fwd|dgrad|wgrad|all] [--one]   (PN2_TC=0 selects the FFMA kernels)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "omni-pq_b200"))
import _pn2 as K  # noqa: E402

SHAPES = [(131072, 128, 256), (131072, 128, 128), (32768, 256, 512), (32768, 256, 256), (32768, 260, 256),
          (8192, 516, 256), (8192, 256, 512), (4096, 256, 256), (1024, 1024, 512), (512, 1024, 512)]


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    shapes = SHAPES[:1] if "--one" in sys.argv else SHAPES
    if os.environ.get("PN2_BENCH_SHAPE"):
        shapes = [SHAPES[int(os.environ["PN2_BENCH_SHAPE"])]]
    dev = "cuda"
    for rows, k, n in shapes:
        yprev = torch.randn(rows, k, device=dev)
        scale, shift = torch.rand(k, device=dev) + 0.5, torch.randn(k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        wt, wp = K.mlp_prep_weights(w, 0, 0, k, n)
        src = K.rows_bnrelu(yprev, rows, k, k, scale, shift)
        gb_f = 4.0 * rows * (k + n) / 1e9
        fl = 2.0 * rows * k * n
        if what in ("fwd", "all"):
            t = timeit(lambda: K.mlp_forward(src, k, n, wt, wp))
            print(f"fwd   {rows:7d} x {k:4d} -> {n:4d}: {t:8.1f} us  {fl / t / 1e6:6.1f} TF  {gb_f / t * 1e6:7.0f} GB/s", flush=True)
        y, _, _ = K.mlp_forward(src, k, n, wt, wp)
        dz = torch.randn(rows, n, device=dev)
        ca, cb, cc = torch.rand(n, device=dev), torch.randn(n, device=dev) * 0.01, torch.randn(n, device=dev) * 0.01
        dy = K.rows_dy(y, dz, rows, n, n, ca, cb, cc)
        if what in ("dgrad", "all"):
            t = timeit(lambda: K.mlp_dgrad_mask(dy, k, wp, yprev, scale, shift, wt=wt))
            gb = 4.0 * rows * (2 * n + 2 * k) / 1e9
            print(f"dgrad {rows:7d} x {n:4d} -> {k:4d}: {t:8.1f} us  {fl / t / 1e6:6.1f} TF  {gb / t * 1e6:7.0f} GB/s", flush=True)
        if what in ("wgrad", "all"):
            t = timeit(lambda: K.mlp_wgrad(dy, src, n, k, 0, 0, dev))
            gb = 4.0 * rows * (2 * n + k) / 1e9
            print(f"wgrad {rows:7d} : {n:4d} x {k:4d}: {t:8.1f} us  {fl / t / 1e6:6.1f} TF  {gb / t * 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
