"""Multi-GPU parity on real hardware: torchrun with one process per GPU over NCCL. Every rank compares
match_sharded / mutual_nn_sharded / extract_sharded with the single-GPU result. Needs >= 2 GPUs (gpurun --gpus 2)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world", [2])
def test_sharded_calls_match_single_gpu_over_nccl(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(REPO / "tests" / "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
