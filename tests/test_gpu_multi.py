"""Multi-GPU (NCCL, one GPU per rank) tests -- skipped on a single-GPU box; run with `gpurun --gpus 2`.

* graphed.GraphedTrainStep with a process group == DistributedDataParallel (gradient averaging, train.py:382);
* the fused SyncBatchNorm path over NCCL == torch.nn.SyncBatchNorm + DDP (models/pq_transformer.py:194)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")]


def _torchrun(script, env=None):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)],
                          capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))


def test_graphed_step_with_nccl_allreduce_matches_ddp(built_lib):
    r = _torchrun("graphed_ddp_worker.py")
    assert r.returncode == 0 and r.stdout.count("GRAPH_DDP_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_fused_syncbn_over_nccl_matches_torch_syncbn(built_lib):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pn2_ref_ext.so")):
        pytest.skip("oracle/_ref/pn2_ref_ext.so (reference kernels) not built")
    r = _torchrun("syncbn_worker.py", {"PN2_SYNCBN_BACKEND": "nccl"})
    assert r.returncode == 0 and r.stdout.count("SYNCBN_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]
    print(r.stdout[-600:])
    # over NVLink peers the exchange must have been served by the fused kernel (4 BatchNorm layers x 2 directions)
    assert r.stdout.count("peer_exchanges=8 []") == 2, r.stdout[-1500:]
