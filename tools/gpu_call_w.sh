#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --tb=short 2>&1 | tail -5 ) > $O/x_pytest.log
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench.json 2> $O/x_bench.err
( PN2_TC_SMALLK_FFMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-gpu ) > $O/x_bench0.json 2> $O/x_bench0.err
echo done
