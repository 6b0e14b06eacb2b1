#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "not ffma" 2>&1 | tail -30 ) > $O/w_pytest.log
( timeout 300 python tools/gemm_bench.py all ) > $O/w_gemm.txt 2>&1
( PN2_BENCH_SHAPE=3 timeout 120 python tools/tile_trace.py fwd ) > $O/w_tile_trace.txt 2>&1
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/w_bench.json 2> $O/w_bench.err
( PN2_TC_WIDE=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/w_bench_narrow.json 2> $O/w_bench_narrow.err
echo done
