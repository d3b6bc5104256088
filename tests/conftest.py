import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def load_golden(name):
    return torch.load(GOLDEN / name, map_location="cpu", weights_only=False)


def batch_from_dict(d):
    from gnn_qot_estimation_b200 import Batch
    kw = {k: v for k, v in d.items() if k != "num_graphs"}
    return Batch(num_graphs=d["num_graphs"], **kw)


# Element-wise parity bar of the prediction tests: |a - b| <= ABS_FLOOR + RTOL * |b| with RTOL = 1e-5 (BASELINE.json
# north_star).  The floor covers entries near zero, where a relative error is meaningless: predictions are O(0.1 .. 1),
# 2e-6 is 1e-5 of the smallest typical magnitude and ~8 fp32 ulps of a typical one (the readout head is a 128-term
# fp32 dot product; the fp32 oracle itself sits 3e-7 from the fp64 one).
ABS_FLOOR = 2e-6


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the 'relative (fp32)' figure of the parity bar."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = max(float(b.abs().max()) if b.numel() else 0.0, 1e-30)
    return float((a - b).abs().max()) / denom if a.numel() else 0.0


def grad_errs(grads, ref, floor_frac=1e-3, exact_zero=()):
    """Per-tensor gradient error for the 1e-5 parity bar: max|a-b| / max(|b|_inf, floor) with
    floor = floor_frac * (largest |gradient entry| over ALL parameters).  Keys in `exact_zero`
    are gradients that vanish in exact arithmetic and exist only as cancellation round-off --
    TransformerConv's lin_key.bias (a constant added to every logit of a softmax row cancels) and
    GATConv's bias in front of a batch-statistics BatchNorm (the mean subtraction removes it); the
    fp64 oracle holds ~1e-16 there, so they are measured against the global gradient scale."""
    scale = max(float(v.detach().abs().max()) for v in ref.values() if v.numel())
    out = {}
    for k, v in ref.items():
        floor = max(scale * (1.0 if k in exact_zero else floor_frac), 1e-30)
        a, b = grads[k].detach().double().cpu(), v.detach().double().cpu()
        assert a.shape == b.shape, k
        out[k] = float((a - b).abs().max()) / max(float(b.abs().max()), floor) if a.numel() else 0.0
    return out


def grad_parity(grads, ref64, ref32, rtol, exact_zero=()):
    """Asserts the gradient parity bar per tensor: error against the fp64 oracle <= rtol, or -- for a
    tensor where fp32 arithmetic itself cannot reach rtol (heavy cancellation, or a ReLU decision
    within round-off of its kink flipping) -- no worse than twice the error the fp32 oracle (the
    reference's own precision and op decomposition) has against fp64."""
    e_ours = grad_errs(grads, ref64, exact_zero=exact_zero)
    e_ref32 = grad_errs(ref32, ref64, exact_zero=exact_zero)
    for k in e_ours:
        assert e_ours[k] <= max(rtol, 2.0 * e_ref32[k]), (k, e_ours[k], e_ref32[k])
    return e_ours
