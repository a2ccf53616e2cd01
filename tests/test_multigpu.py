"""Multi-GPU parity: tests/mgpu_worker.py under torchrun, one rank per GPU, against the oracle.  Skips on a
box with fewer than two GPUs (run it with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_paths_vs_oracle(exchange):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n, 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, LAT_EXCHANGE=exchange, LAT_SPIN_TIMEOUT_MS="20000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-6000:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):  # keep the workers' output where a gpurun call brings it back
        with open(os.path.join(out_dir, f"mgpu_worker_{exchange}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, "mgpu_worker failed"
    assert "MISMATCH" not in r.stdout
