"""GPU, N >= 2: sharded top-K over real NCCL equals the single-GPU result (tools/multi_gpu_check.py under torchrun).
Skipped on single-GPU boxes; the host-side logic is covered on CPU by tests/test_sharded_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_equals_single_gpu_over_nccl():
    n = min(torch.cuda.device_count(), 8)
    with socket.socket() as s:                       # a free rendezvous port (concurrent runs must not collide)
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded == single-GPU == oracle: True" in r.stdout
    assert "sharded == single-GPU: True" in r.stdout and "sharded == single-GPU: False" not in r.stdout
    assert "query-sharded == single-GPU == oracle: True" in r.stdout
    assert "postings pool-sharded (NCCL) == single-GPU: True" in r.stdout
    # fused NVLink exchange (symmetric memory): must agree with the NCCL path; if symmetric memory is unavailable on this
    # box the tool says so explicitly and the fused kernels were NOT exercised here -> skip instead of a silent pass
    assert "== NCCL path: False" not in r.stdout
    if "fused P2P exchange unavailable" in r.stdout:
        pytest.skip("symmetric memory unavailable: fused exchange not exercised (NCCL path verified)")
    assert "jaccard fused P2P exchange == NCCL path: True" in r.stdout
    assert "postings fused P2P exchange == NCCL path: True" in r.stdout
    assert "fused P2P == NCCL path: True" in r.stdout
