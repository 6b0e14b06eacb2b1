#!/bin/bash
# round-2 GPU call B: full GPU suite with the new parity tests, new bench line, FFMA-path subprocess test x10
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > $O/b_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 5 --extras ) > $O/b_bench.json 2> $O/b_bench.err
for i in 1 2 3 4 5 6 7 8 9 10; do
  ( PN2_TC=0 timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "gemm or transposes or config1 or three_layer or without_features or fp_matches" 2>&1 | tail -3 ) >> $O/b_ffma_loop.txt
done
echo done
