"""Where does a tensor-core GEMM CTA spend its time?  Installs the per-CTA phase trace (pn2_debug_gemm_trace) and
prints, per shape: mean prologue / main-loop / epilogue time per CTA, the gap between consecutive CTAs on an SM,
and the kernel span (development aid; run on the B200 box).
    python tools/gemm_trace.py [fwd|dgrad|wgrad|all]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "omni-pq_b200"))
import _pn2 as K  # noqa: E402

SHAPES = [(131072, 128, 256), (131072, 128, 128), (32768, 256, 512), (32768, 256, 256), (8192, 516, 256), (1024, 1024, 512)]
K.lib.pn2_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
CAP = 1 << 15


def traced(fn, label):
    fn()
    torch.cuda.synchronize()
    buf = torch.zeros(CAP, 6, dtype=torch.int64, device="cuda")
    K.lib.pn2_debug_gemm_trace(buf.data_ptr(), CAP)
    fn()
    torch.cuda.synchronize()
    K.lib.pn2_debug_gemm_trace(None, 0)
    t = buf.cpu()
    t = t[t[:, 1] > 0]
    n = t.shape[0]
    sm, t0, t1, t2, t3, kb = (t[:, i].double() for i in range(6))
    span = (t3.max() - t0.min()) / 1e3
    pro, main, epi = (t1 - t0).mean() / 1e3, (t2 - t1).mean() / 1e3, (t3 - t2).mean() / 1e3
    gaps, busy = [], []
    for s in sm.unique():
        sel = sm == s
        a, b = t0[sel].sort().values, t3[sel].sort().values
        busy.append(float((t3[sel] - t0[sel]).sum()) / 1e3)
        if a.numel() > 1:
            gaps.append(float((a[1:] - b[:-1]).clamp_min(0).mean()) / 1e3)
    print(f"{label}: {n} CTAs on {sm.unique().numel()} SMs, k-blocks {int(kb.max())}, span {span:.1f} us | per CTA: "
          f"prologue {pro:.2f} main {main:.2f} ({main / max(kb.mean(), 1):.2f}/k-block) epilogue {epi:.2f} us | "
          f"gap between CTAs on an SM {sum(gaps) / max(len(gaps), 1):.2f} us | SM busy {sum(busy) / len(busy):.1f} us", flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda"
    for rows, k, n in SHAPES:
        yprev = torch.randn(rows, k, device=dev)
        scale, shift = torch.rand(k, device=dev) + 0.5, torch.randn(k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        wt, wp = K.mlp_prep_weights(w, 0, 0, k, n)
        src = K.rows_bnrelu(yprev, rows, k, k, scale, shift)
        if what in ("fwd", "all"):
            traced(lambda: K.mlp_forward(src, k, n, wt, wp), f"fwd   {rows} x {k} -> {n}")
        y, _, _ = K.mlp_forward(src, k, n, wt, wp)
        dz = torch.randn(rows, n, device=dev)
        ca, cb, cc = torch.rand(n, device=dev), torch.randn(n, device=dev) * 0.01, torch.randn(n, device=dev) * 0.01
        dy = K.rows_dy(y, dz, rows, n, n, ca, cb, cc)
        if what in ("dgrad", "all"):
            traced(lambda: K.mlp_dgrad_mask(dy, k, wp, yprev, scale, shift, wt=wt), f"dgrad {rows} x {n} -> {k}")
        if what in ("wgrad", "all"):
            traced(lambda: K.mlp_wgrad(dy, src, n, k, 0, 0, dev), f"wgrad {rows} : {n} x {k}")


if __name__ == "__main__":
    main()
