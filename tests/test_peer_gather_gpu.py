"""`dl_peer_allgather` (all-gather as one kernel over NVLink peer memory) needs >= 2 GPUs: runs
tools/test_peer_gather.py under torchrun (correctness against NCCL over 30 mixed-size calls, halo
exchange, CUDA-graph replay with an odd number of calls per replay); skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with NVLink peer access")
def test_peer_allgather_two_ranks():
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
                        os.path.join(ROOT, "tools", "test_peer_gather.py")],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "RESULT True True True" in p.stdout
