#!/bin/bash
# geometry prefetch: parity test + bench with / without
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q --tb=short -k "graphed" 2>&1 | tail -30 ) > $O/r_pytest.log
( timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/r_bench.json 2> $O/r_bench.err
( timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-prefetch ) > $O/r_bench_nopf.json 2> $O/r_bench_nopf.err
echo done
