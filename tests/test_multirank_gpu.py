"""GPU, multi-rank (N > 1 correctness, not arithmetic on paper): spawns one process per GPU with
torch.distributed.run when at least 2 GPUs are visible (`gpurun --gpus 2`); skipped on a 1-GPU box.  The worker
(tests/mp/peer_worker.py) checks the library's peer-memory all-reduce against NCCL (bitwise identical results on every
rank, eager and captured) and the data-parallel captured train step against the mean of the per-shard gradients."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2])
def test_peer_allreduce_and_dp_step_multi_rank(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world}); {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mp", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "OK world" in r.stdout
