#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
for sp in 0 16; do
( PN2_WGRAD_SPARE=$sp timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/s_bench_wspare$sp.json 2> $O/s_bench_wspare$sp.err
done
echo done
