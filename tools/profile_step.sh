#!/bin/bash
# ncu evidence for one bench step on the B200 box (run under gpurun).  Usage: tools/profile_step.sh <tag>
# Writes (all small enough to be merged back): gpurun_out/<tag>_launches.csv  (every launch, gpu__time_duration),
# <tag>_gemm_raw.csv / <tag>_rest_raw.csv (ncu --set full, --page raw).
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-graph"
# eager steps (no graph): warm-up/counting step + 1 warm-up + 1 timed + e2e(1+1) + 3 profiled = 8 steps
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $OUT/${TAG}_launches.csv $B > /dev/null 2>&1
# one step's worth of GEMM launches (47 per step; skip the first two steps)
timeout 420 ncu --set full --clock-control none -k regex:"gemm_tc|wgrad_tc" -s 94 -c 47 -o /tmp/${TAG}_gemm $B > /dev/null 2>&1
ncu -i /tmp/${TAG}_gemm.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_raw.csv 2>/dev/null
if [ "${PROFILE_REST:-0}" = "1" ]; then  # the non-GEMM kernels (unchanged since profiles/r2mid_ncu_rest.json)
  timeout 300 ncu --set full --clock-control none -k regex:"fps_multipick|ball_query|bn_relu_pool|pool_bwd|fp_interpolate_kernel" -s 18 -c 18 -o /tmp/${TAG}_rest $B > /dev/null 2>&1
  ncu -i /tmp/${TAG}_rest.ncu-rep --page raw --csv > $OUT/${TAG}_rest_raw.csv 2>/dev/null
fi
ls -la $OUT | tail -8
