#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), the GPU suite, the bench line (ours and the reference arm)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/v_smoke.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 ) > $O/v_pytest.log
( timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 --extras ) > $O/v_bench.json 2> $O/v_bench.err
( timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > $O/v_bench_ref.json 2> $O/v_bench_ref.err
echo done
