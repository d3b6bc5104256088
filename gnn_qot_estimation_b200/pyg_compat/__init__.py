"""Minimal ``torch_geometric`` stand-in for the names the reference imports, so its scripts run in
an environment without PyTorch Geometric:

    topological_training/models.py:3   torch_geometric.nn: global_mean_pool, TransformerConv, NNConv
    lightpath_training/models.py:3     torch_geometric.nn: GATConv, BatchNorm
    */train.py:2, */test.py:2|4        torch_geometric.loader.DataLoader
    */dataset.py:6                     torch_geometric.utils.from_networkx

``install()`` registers these modules under the ``torch_geometric`` name in ``sys.modules`` (it
refuses to shadow a real installation).  The layers are the B200 kernels
(:mod:`gnn_qot_estimation_b200.nn`); ``Data`` / ``Batch`` / ``DataLoader`` / ``from_networkx``
restate PyG's host-side behaviour that the reference relies on.
"""
from __future__ import annotations

import importlib.util
import sys
import types

from .data import Batch, Data  # noqa: F401
from .loader import DataLoader  # noqa: F401
from .utils import from_networkx  # noqa: F401


def install(force: bool = False) -> None:
    if not force and "torch_geometric" not in sys.modules and importlib.util.find_spec("torch_geometric") is not None:
        raise RuntimeError("a real torch_geometric is installed; pyg_compat.install() would shadow it "
                           "(pass force=True to do so anyway)")
    from .. import nn as qnn
    root = types.ModuleType("torch_geometric")
    root.__path__ = []                                   # mark as a package
    sub = {
        "nn": {k: getattr(qnn, k) for k in ("TransformerConv", "NNConv", "GATConv", "BatchNorm", "global_mean_pool")},
        "data": {"Data": Data, "Batch": Batch},
        "loader": {"DataLoader": DataLoader},
        "utils": {"from_networkx": from_networkx},
    }
    sys.modules["torch_geometric"] = root
    for name, members in sub.items():
        m = types.ModuleType(f"torch_geometric.{name}")
        for k, v in members.items():
            setattr(m, k, v)
        setattr(root, name, m)
        sys.modules[f"torch_geometric.{name}"] = m
