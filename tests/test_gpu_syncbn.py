"""SyncBatchNorm semantics of the fused path (models/pq_transformer.py:194 converts every hot-path
BatchNorm): world_size 2 on ONE GPU over gloo, ours vs torch.nn.SyncBatchNorm under DDP on the reference's own CUDA
kernels.  Covers the device-side row count (no host sync), unequal rows per rank, and rank-local dgamma/dbeta
(the all-reduced sums must only enter the input-gradient coefficients)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fused_syncbn_matches_torch_syncbn_under_ddp(built_lib):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pn2_ref_ext.so")):
        pytest.skip("oracle/_ref/pn2_ref_ext.so (reference kernels) not built")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "syncbn_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("SYNCBN_OK") == 2, r.stdout[-2000:]
