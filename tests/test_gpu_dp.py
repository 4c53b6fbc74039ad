"""Two-GPU checks of the peer-memory data-parallel step (needs >= 2 devices; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("provider", ["auto", "ipc"])   # in-switch reduce / broadcast (NVLS) where available; peer loads / stores
def test_distributed_fused_adam_and_occupancy_replicas_on_two_gpus(provider):
    port = 29700 + (os.getpid() % 200) + (7 if provider == "ipc" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, CEDNERF_DP_PROVIDER=provider))
    assert out.returncode == 0 and "DP-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-6000:]
