"""Per-tile timeline of the persistent async GEMM (pn2_debug_gemm_trace2): where does a tile's time go?
    PN2_BENCH_SHAPE=3 python tools/tile_trace.py [fwd|dgrad]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "omni-pq_b200"))
import _pn2 as K  # noqa: E402

SHAPES = [(131072, 128, 256), (131072, 128, 128), (32768, 256, 512), (32768, 256, 256), (32768, 260, 256),
          (8192, 516, 256), (8192, 256, 512), (4096, 256, 256), (1024, 1024, 512), (512, 1024, 512)]
K.lib.pn2_debug_gemm_trace2.argtypes = [ctypes.c_void_p, ctypes.c_int]
NAMES = ["start", "kb0 staged", "last kb staged", "acc complete", "epilogue done", "mma: A seen", "mma: B seen", "loader: first B"]


def run(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ctas = 148
    buf = torch.zeros(ctas, 8, 8, dtype=torch.int64, device="cuda")
    K.lib.pn2_debug_gemm_trace2(buf.data_ptr(), ctas)
    fn()
    torch.cuda.synchronize()
    K.lib.pn2_debug_gemm_trace2(None, 0)
    t = buf.cpu().double()
    print(label)
    t0 = t[t[:, 0, 0] > 0][:, 0, 0].min()
    for ti in range(5):
        sel = t[:, ti, 0] > 0
        if sel.sum() == 0:
            break
        x = t[sel, ti]
        rel = (x - x[:, :1]) / 1e3  # us since this tile's start
        start = (x[:, 0] - t0) / 1e3
        row = "  ".join(f"{NAMES[k]} {rel[:, k].mean():6.2f}" for k in (1, 2, 3, 4, 5, 6, 7))
        print(f"  tile #{ti} ({int(sel.sum())} CTAs, starts at {start.mean():6.2f} us): {row}")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    rows, k, n = SHAPES[int(os.environ.get("PN2_BENCH_SHAPE", "3"))]
    dev = "cuda"
    yprev = torch.randn(rows, k, device=dev)
    scale, shift = torch.rand(k, device=dev) + 0.5, torch.randn(k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    wt, wp = K.mlp_prep_weights(w, 0, 0, k, n)
    src = K.rows_bnrelu(yprev, rows, k, k, scale, shift)
    if what == "fwd":
        run(lambda: K.mlp_forward(src, k, n, wt, wp), f"fwd {rows} x {k} -> {n}")
    else:
        y, _, _ = K.mlp_forward(src, k, n, wt, wp)
        dz = torch.randn(rows, n, device=dev)
        ca, cb, cc = torch.rand(n, device=dev), torch.randn(n, device=dev) * 0.01, torch.randn(n, device=dev) * 0.01
        dy = K.rows_dy(y, dz, rows, n, n, ca, cb, cc)
        run(lambda: K.mlp_dgrad_mask(dy, k, wp, yprev, scale, shift, wt=wt), f"dgrad {rows} x {n} -> {k}")


if __name__ == "__main__":
    main()
