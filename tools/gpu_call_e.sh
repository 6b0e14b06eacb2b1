#!/bin/bash
# round-2 GPU call E: per-tile timeline of the async kernel; arbiter numbers with the rounded lo half
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for sh in 3 1 8; do
  ( PN2_BENCH_SHAPE=$sh timeout 120 python tools/tile_trace.py fwd ) >> $O/e_tile_trace.txt 2>&1
  ( PN2_BENCH_SHAPE=$sh timeout 120 python tools/tile_trace.py dgrad ) >> $O/e_tile_trace.txt 2>&1
done
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q --tb=short -k "not ffma" 2>&1 | tail -80 ) > $O/e_pytest.log
echo done
