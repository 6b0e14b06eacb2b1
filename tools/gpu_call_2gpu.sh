#!/bin/bash
# 2-GPU call (strictly time-boxed): the NCCL tests (fused SyncBatchNorm exchange over peer memory, graph + NCCL == DDP)
# and the 2-rank bench line
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -s 2>&1 | tail -60 ) > $O/t_pytest_multi.log
( timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/t_bench_2gpu.json 2> $O/t_bench_2gpu.err
echo done
