"""One process per GPU over NCCL + CUDA IPC (tests/mp_worker.py). Needs >= 2 devices; on a single-GPU box the test is
skipped (the same protocol is covered on one device by tests/test_gpu_heat_halo.py and on CPU by
tests/test_multi_rank_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_paths_against_unsharded_oracle(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"MP_WORKER_OK {world}" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
