"""Measured TF32 GEMM peak of this GPU (the tensor-roofline denominator of the 3xTF32 shared-MLP
GEMMs must be a *measured TF32* peak, MEASURED_PEAKS.json only holds bf16).  cuBLAS fp32 matmul with TF32 allowed,
8192^3: best of 10 (burst, for a kernel timed alone) and back to back for ~3 s (sustained, for a kernel inside a long
step).  bench.py calls measure() in its setup; run as a script it prints one JSON line."""
import json
import time


def measure(n=8192, sustain_s=3.0):
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        best = 0.0
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(a, b, out=c)
            e.record()
            torch.cuda.synchronize()
            best = max(best, flops / (s.elapsed_time(e) * 1e-3) / 1e12)
        # sustained: back-to-back launches for sustain_s seconds, one event pair around the lot
        reps = max(10, int(sustain_s / (flops / (best * 1e12))))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e.record()
        torch.cuda.synchronize()
        sustained = reps * flops / (s.elapsed_time(e) * 1e-3) / 1e12
        return {"tf32_tflops": best, "tf32_tflops_sustained": sustained, "n": n, "reps_sustained": reps,
                "how": "torch.matmul fp32 with allow_tf32=True (cuBLAS TF32 tensor-core GEMM), 8192^3, 2*N^3 flops: best of "
                       "10 (burst) and back to back (sustained), CUDA events",
                "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


if __name__ == "__main__":
    print(json.dumps(measure()))
