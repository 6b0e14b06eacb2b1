#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --tb=short 2>&1 | tail -30 ) > $O/x_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench.json 2> $O/x_bench.err
( PN2_WGRAD_STREAM_ROWS=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench_ws0.json 2> $O/x_bench_ws0.err
( PN2_WGRAD_STREAM_ROWS=100000000 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench_wsall.json 2> $O/x_bench_wsall.err
echo done
