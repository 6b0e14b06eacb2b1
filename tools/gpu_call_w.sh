#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 300 python tools/gemm_bench.py fwd ) > $O/x_gemm.txt 2>&1
( timeout 300 python tools/gemm_bench.py dgrad ) >> $O/x_gemm.txt 2>&1
( PN2_TC_STAGGER=0 timeout 300 python tools/gemm_bench.py fwd ) > $O/x_gemm0.txt 2>&1
( PN2_TC_STAGGER=0 timeout 300 python tools/gemm_bench.py dgrad ) >> $O/x_gemm0.txt 2>&1
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench.json 2> $O/x_bench.err
( PN2_TC_STAGGER=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench_st0.json 2> $O/x_bench_st0.err
echo done
