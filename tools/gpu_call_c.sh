#!/bin/bash
# round-2 GPU call C: async TS GEMM kernel vs round-1 pbulk kernel (micro-bench, tests, step bench)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 300 python tools/gemm_bench.py fwd ) > $O/c_gemm_async_fwd.txt 2>&1
( timeout 300 python tools/gemm_bench.py dgrad ) > $O/c_gemm_async_dgrad.txt 2>&1
( PN2_TC_ASYNC=0 timeout 300 python tools/gemm_bench.py fwd ) > $O/c_gemm_pbulk_fwd.txt 2>&1
( PN2_TC_ASYNC=0 timeout 300 python tools/gemm_bench.py dgrad ) > $O/c_gemm_pbulk_dgrad.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 ) > $O/c_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/c_bench.json 2> $O/c_bench.err
( PN2_TC_ASYNC=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/c_bench_pbulk.json 2> $O/c_bench_pbulk.err
echo done
