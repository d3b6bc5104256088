"""GPU, >= 2 devices: over NCCL, averaged per-rank gradients of the B200 TopologicalGNN equal the
gradients of the concatenated batch, and the replicas stay bit-identical after CUDA-graphed DDP
steps (scripts/ddp_check.py under torchrun, one process per GPU).  Skipped on a 1-GPU box; the
host-side logic is covered by tests/test_distributed_cpu.py (gloo, world 2)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("exchange", ["graph", "eager"])
def test_ddp_over_nccl(exchange):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, QOT_DDP_EXCHANGE=exchange, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "scripts" / "ddp_check.py")],
                         capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "replicas identical after graphed steps: True" in out.stdout
