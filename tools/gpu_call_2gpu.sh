#!/bin/bash
# 2-GPU call (strictly time-boxed): bench at N=2, then the NCCL tests
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 ) > $O/q_bench_2gpu.json 2> $O/q_bench_2gpu.err
( timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -x 2>&1 | tail -40 ) > $O/q_pytest_multi.log
echo done
