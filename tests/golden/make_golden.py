"""Freeze golden vectors for the index ops (FPS order, ball-query idx, 3-NN idx / dist2).

    python tests/golden/make_golden.py --backend oracle     # CPU restatement (runs anywhere)
    python tests/golden/make_golden.py --backend ref        # the REFERENCE's own CUDA kernels
                                                            # (oracle/_ref/pn2_ref_ext.so, needs a GPU)

The reference pins no known-answer vectors for this path, and its kernels only run on a
GPU.  So `pn2_golden_ref.npz` is generated ON THE B200 BOX from the unmodified reference kernels
compiled by oracle/build_ref.py (written to gpurun_out/ there, then committed here), and
`pn2_golden_oracle.npz` from the CPU oracle in the build container; tests/test_oracle_cpu.py checks the
oracle against both, tests/test_gpu_ops.py checks the B200 kernels against both.
Each file stores, per case of tests/cases.py, every index tensor plus a digest of the input cloud.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["oracle", "ref"], default="oracle")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.backend == "oracle":
        from oracle import pn2_oracle as O
        ext, device = O.ext, "cpu"
    else:
        from oracle import build_ref
        ext, device = build_ref.load(), "cuda"
        assert ext is not None, "oracle/_ref/pn2_ref_ext.so missing: run oracle/build_ref.py first"
    blob = {}
    for name in cases.CASES:
        outs, dig = cases.run_case(name, ext, device)
        blob[f"{name}/input_digest"] = np.frombuffer(dig.encode(), dtype=np.uint8)
        for k, v in outs.items():
            blob[f"{name}/{k}"] = v.numpy()
        print(name, dig, {k: tuple(v.shape) for k, v in outs.items()})
    out = args.out or os.path.join(HERE, f"pn2_golden_{args.backend}.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
