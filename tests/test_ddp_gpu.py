"""GPU, >= 2 devices: over NCCL / NVLink peer memory, averaged per-rank gradients of the B200 TopologicalGNN
equal the gradients of the concatenated batch, and the replicas stay bit-identical after CUDA-graphed DDP
steps (scripts/ddp_check.py under torchrun, one process per GPU).  Skipped on a 1-GPU box; the
host-side logic is covered by tests/test_distributed_cpu.py (gloo, world 2).

exchange = "peer": gradient exchange + SGD in one native launch (one-shot all-reduce over peer memory,
csrc/ddp_step.cu); "eager": flat NCCL all-reduce between two captured graphs.  (The third mode, "graph" --
the NCCL call captured inside the step graph -- is exercised by bench.py's cfg 3 block, not here.)"""
import os
import signal
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("exchange", ["peer", "eager"])
def test_ddp_over_nvlink(exchange):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, QOT_DDP_EXCHANGE=exchange, MASTER_ADDR="127.0.0.1")
    port = 29533 + (0 if exchange == "peer" else 1)
    proc = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                             "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "scripts" / "ddp_check.py")],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=str(ROOT),
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=150)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)                   # the whole torchrun group: a hung rank must not outlive the test
        out, err = proc.communicate()
        pytest.fail(f"ddp_check ({exchange}) did not finish in 150 s\n{out[-1500:]}\n{err[-1500:]}")
    assert proc.returncode == 0, (out[-2000:], err[-3000:])
    assert "replicas identical after graphed steps: True" in out
    assert f"exchange={exchange}" in out, out[-1500:]          # no silent fallback to another exchange
    assert "solo trainer on rank 0 stepped without a collective: True" in out


def test_second_device_in_one_process():
    """One process driving two GPUs (not the deployment shape -- one process per GPU is -- but it must work):
    per-device kernel attributes (> 48 KB dynamic shared memory), workspaces and streams follow the tensors' device,
    not the current one.  LightpathGNN eval and a TopologicalGNN train step on cuda:1 while cuda:0 is current."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import sys
    sys.path.insert(0, str(ROOT))
    from gnn_qot_estimation_b200 import LightpathGNN, TopologicalGNN, synthetic
    torch.cuda.set_device(0)
    outs = {}
    for d in (0, 1):
        dev = torch.device("cuda", d)
        torch.manual_seed(0)
        m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0).to(dev).eval()
        b = synthetic.lightpath_store(200, seed=4).to(dev).collate(range(0, 200))
        with torch.no_grad():
            o, lb = m(b)
        t = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(dev)
        tb = synthetic.nsfnet_store(32, seed=1).to(dev).collate(range(0, 32))
        loss = torch.nn.SmoothL1Loss()(t(tb), tb.y.view(-1, 3))
        loss.backward()
        outs[d] = (o.cpu(), lb.cpu(), loss.detach().cpu(), t.conv1.lin_query.weight.grad.cpu())
        assert o.device == dev and torch.cuda.current_device() == 0
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)                               # same bits on both devices
