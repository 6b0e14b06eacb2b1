#!/bin/bash
# round-2 GPU call G: converged MMA warps (async fwd/dgrad + wgrad kernels), ticket draw off the critical path
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 300 python tools/gemm_bench.py all ) > $O/n_gemm.txt 2>&1
for sh in 3 1; do
  ( PN2_BENCH_SHAPE=$sh timeout 120 python tools/tile_trace.py fwd ) >> $O/n_tile_trace.txt 2>&1
  ( PN2_BENCH_SHAPE=$sh timeout 120 python tools/tile_trace.py dgrad ) >> $O/n_tile_trace.txt 2>&1
done
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py -m gpu -q --tb=short -k "not ffma" 2>&1 | tail -40 ) > $O/n_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/n_bench.json 2> $O/n_bench.err
echo done
