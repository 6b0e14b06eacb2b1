#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-graph"
# the 11 GEMM launches of the step that the 420 s budget of tools/profile_step.sh did not reach (launches 37..47)
timeout 400 ncu --set full --clock-control none -k regex:"gemm_tc|wgrad_tc" -s 130 -c 11 -o /tmp/r2b_gemm $B > /dev/null 2>&1
ncu -i /tmp/r2b_gemm.ncu-rep --page raw --csv > $O/r2b_gemm_raw.csv 2>/dev/null
( timeout 300 python tools/gemm_bench.py all ) > $O/y_gemm.txt 2>&1
rm -f $O/y_trace.txt
for sh in 0 1 3; do for w in fwd dgrad; do ( PN2_BENCH_SHAPE=$sh timeout 120 python tools/tile_trace.py $w ) >> $O/y_trace.txt 2>&1; done; done
echo done
