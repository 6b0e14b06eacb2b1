#!/bin/bash
# round-2 GPU call D: async kernel after the context fix (tests + ncu source profile), FFMA-path flake loop with the
# self-diagnosing test, selective PDL A/B
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py -m gpu -q --tb=short -k "not ffma" 2>&1 | tail -80 ) > $O/d_pytest.log
( PN2_BENCH_SHAPE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_async -s 3 -c 1 -f -o $O/d_async_fwd python tools/gemm_bench.py fwd ) > $O/d_ncu_fwd.log 2>&1
( PN2_BENCH_SHAPE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_async -s 3 -c 1 -f -o $O/d_async_dgrad python tools/gemm_bench.py dgrad ) > $O/d_ncu_dgrad.log 2>&1
for i in $(seq 1 16); do
  ( PN2_TC=0 timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --tb=short -k "gemm or transposes or config1 or three_layer or without_features or fp_matches" 2>&1 | grep -E "passed|failed|ours vs" ) >> $O/d_ffma_loop.txt
done
for i in $(seq 1 12); do
  ( PN2_TC=0 CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --tb=short -k "gemm or transposes or config1 or three_layer or without_features or fp_matches" 2>&1 | grep -E "passed|failed|ours vs" ) >> $O/d_ffma_loop_blocking.txt
done
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/d_bench.json 2> $O/d_bench.err
( PN2_PDL_SMALL=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/d_bench_pdlsmall.json 2> $O/d_bench_pdlsmall.err
( PN2_TC_ASYNC=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/d_bench_pbulk.json 2> $O/d_bench_pbulk.err
echo done
