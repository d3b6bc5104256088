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


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the 'relative (fp32)' figure of the parity bar."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = max(float(b.abs().max()) if b.numel() else 0.0, 1e-30)
    return float((a - b).abs().max()) / denom if a.numel() else 0.0
