#!/bin/bash
# 2-GPU call (strictly time-boxed): the NCCL tests (fused SyncBatchNorm exchange over peer memory, graph + NCCL == DDP)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -s 2>&1 | tail -60 ) > $O/t_pytest_multi.log
echo done
