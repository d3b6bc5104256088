"""CPU tests of the torch_geometric stand-in (gnn_qot_estimation_b200.pyg_compat): from_networkx
ordering, Data semantics the reference datasets rely on, Batch.from_data_list == oracle.collate_ref,
DataLoader batching, and sys.modules installation."""
import sys

import networkx as nx
import pytest
import torch

from gnn_qot_estimation_b200 import pyg_compat
from gnn_qot_estimation_b200.pyg_compat import Batch, Data, DataLoader, from_networkx
from oracle import collate_ref


def _graph(n, seed):
    g = nx.gnm_random_graph(n, 2 * n, seed=seed)
    for u, v in g.edges():
        g[u][v]["freq"] = 0.1 * (u + v)
        g[u][v]["num_spans"] = float(u * v)
    g.graph["labels"] = {"osnr": 20.0, "snr": 18.0, "ber": 1e-4}
    return g


def test_from_networkx_edge_order_and_attrs():
    g = _graph(9, 0)
    g.add_edge(3, 3, freq=0.5, num_spans=1.0)                      # a self loop appears once
    d = from_networkx(g)
    exp = list(g.to_directed().edges())
    assert d.edge_index.dtype == torch.int64 and d.edge_index.t().tolist() == [list(e) for e in exp]
    assert d.num_nodes == 9
    assert torch.allclose(d.freq, torch.tensor([g[u][v]["freq"] for u, v in exp], dtype=d.freq.dtype))
    # relabelling of non-integer node names in G.nodes() order
    h = nx.Graph()
    h.add_edges_from([("b", "a"), ("a", "c")])
    assert from_networkx(h).edge_index.t().tolist() == [[0, 1], [1, 0], [1, 2], [2, 1]]


def test_data_none_assignment_removes_key():
    d = Data(x=torch.zeros(3, 2), edge_index=torch.zeros(2, 4, dtype=torch.int64))
    assert d.num_nodes == 3
    d.x = None                                                     # topological_training/dataset.py:107
    assert d.x is None and "x" not in d
    d.node_ids = torch.arange(3)
    assert d.num_nodes == 3
    with pytest.raises(AttributeError):
        d.not_there


def _reference_style_item(g):
    """What TopologicalDataset.__getitem__ builds (dataset.py:75-123), minus scaling."""
    d = from_networkx(g)
    d.node_ids = torch.arange(d.num_nodes)
    d.edge_attr = torch.rand(d.edge_index.shape[1], 4)
    d.x = None
    d.y = torch.rand(3)
    return d


def test_batch_from_data_list_equals_collate_ref():
    items = [_reference_style_item(_graph(n, n)) for n in (5, 9, 7)]
    b = Batch.from_data_list(items)
    ref = collate_ref([{"num_nodes": d.num_nodes, "edge_index": d.edge_index, "edge_attr": d.edge_attr,
                        "node_ids": d.node_ids, "y": d.y} for d in items])
    assert b.num_graphs == 3 and b.x is None
    for k in ("edge_index", "edge_attr", "node_ids", "batch", "ptr"):
        assert torch.equal(getattr(b, k), getattr(ref, k)), k
    assert torch.equal(b.y, ref.y) and b.y.shape == (9,)            # train.py:112 views it as [-1, 3]
    assert torch.equal(b.edge_ptr, ref.edge_ptr)
    assert b.to("cpu").num_graphs == 3


def test_dataloader_batches_like_the_reference_loops():
    items = [_reference_style_item(_graph(6, s)) for s in range(10)]
    batches = list(DataLoader(items, batch_size=4, shuffle=False))
    assert [b.num_graphs for b in batches] == [4, 4, 2]
    assert sum(b.num_nodes for b in batches) == 60
    # an oracle model consumes the batch exactly as the reference model would (forward(data))
    from oracle import TopologicalGNNOracle
    out = TopologicalGNNOracle(6, 16, 3, 4, dropout_p=0.0)(batches[0])
    assert out.shape == (4, 3)


def test_install_registers_modules():
    if "torch_geometric" in sys.modules and not isinstance(sys.modules["torch_geometric"].__dict__.get("__path__"), list):
        pytest.skip("real torch_geometric present")
    pyg_compat.install()
    try:
        from torch_geometric.loader import DataLoader as DL
        from torch_geometric.nn import BatchNorm, GATConv, NNConv, TransformerConv, global_mean_pool
        from torch_geometric.utils import from_networkx as fnx
        assert DL is DataLoader and fnx is from_networkx and callable(global_mean_pool)
        assert TransformerConv.__module__ == "gnn_qot_estimation_b200.nn" and GATConv and NNConv and BatchNorm
    finally:
        for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
            del sys.modules[k]
