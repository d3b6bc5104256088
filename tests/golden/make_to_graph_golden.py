"""Generates tests/golden/to_graph_lightpath.pt by running the REFERENCE's own code, unchanged:
/root/reference/to_graph.py::create_lightpath_graph (over an in-memory stand-in for xarray, the only
absent import; the stand-in serves the arrays of synthetic.network_status_samples) -> pickle ->
/root/reference/lightpath_training/dataset.py::LightpathDataset (over the torch_geometric stand-in).
Run in the build container only (needs /root/reference):  python tests/golden/make_to_graph_golden.py"""
import io
import contextlib
import os
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path(os.environ.get("QOT_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
FEATURES = ["mod_order", "path_len", "num_spans", "freq"]          # store_graphs.py:51-56

# (num_samples, num_links, num_freqs, seed, spacing): the last set sits exactly on the 0.05 threshold
# of to_graph.py:300 (adjacent channels interact or not by float64 rounding alone)
CASES = [(6, 12, 64, 0, 0.0375), (4, 8, 48, 1, 0.02), (4, 10, 72, 2, 0.05), (3, 40, 64, 3, 0.0375)]


class _Var:
    def __init__(self, a):
        self.values = a

    def isel(self, sample):
        return _Var(self.values[sample])


class _Dataset(dict):
    def close(self):
        pass


def install_fake_xarray(samples):
    """xarray.open_dataset(path) -> the synthetic arrays, whatever the path."""
    ds = _Dataset(data=_Var(samples["data"]), target=_Var(samples["target"]),
                  lp_feat=_Var(np.array(samples["lp_feat"])), metric=_Var(np.array(samples["metric"])),
                  link=_Var(np.arange(samples["data"].shape[2])), freq=_Var(samples["freqs"]),
                  sample=_Var(np.arange(samples["data"].shape[0])))
    mod = types.ModuleType("xarray")
    mod.open_dataset = lambda path: ds
    sys.modules["xarray"] = mod


def reference_graphs(samples, freq_threshold=0.05, representation="lightpath"):
    """[networkx.Graph] from the reference's create_lightpath_graph / create_topological_graph, one per sample."""
    install_fake_xarray(samples)
    sys.path.insert(0, str(REF))
    sys.modules.pop("to_graph", None)
    import to_graph                                                   # the reference's file
    to_graph.dataset_metadata_cache = None
    out = []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(samples["data"].shape[0]):
            if representation == "lightpath":
                out.append(to_graph.create_lightpath_graph(i, FEATURES, "unused.nc", freq_threshold=freq_threshold))
            else:
                out.append(to_graph.create_topological_graph(i, FEATURES, "unused.nc"))
    sys.path.remove(str(REF))
    sys.modules.pop("to_graph", None)
    sys.modules.pop("xarray", None)
    return out


def reference_data_objects(graphs):
    """The reference's LightpathDataset over pickles of those graphs -> [Data]."""
    from gnn_qot_estimation_b200 import pyg_compat
    pyg_compat.install()
    sys.path.insert(0, str(REF))
    try:
        from lightpath_training.dataset import LightpathDataset          # the reference's file
        with tempfile.TemporaryDirectory() as d:
            for i, g in enumerate(graphs):
                with open(os.path.join(d, f"graph_{i:04d}.gpickle"), "wb") as f:
                    pickle.dump(g, f)
            ds = LightpathDataset(directory=d)
            return [ds[i] for i in range(len(ds))], list(ds.node_features)
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k.split(".")[0] in ("lightpath_training", "constants")]:
            del sys.modules[k]


def reference_topological_data_objects(graphs):
    """The reference's TopologicalDataset over pickles of those graphs -> [Data] (edge ORDER included:
    edges are added in ascending conn_id order, to_graph.py:147-178, so it is well defined)."""
    from gnn_qot_estimation_b200 import pyg_compat
    pyg_compat.install()
    sys.path.insert(0, str(REF))
    try:
        from topological_training.dataset import TopologicalDataset      # the reference's file
        with tempfile.TemporaryDirectory() as d:
            for i, g in enumerate(graphs):
                with open(os.path.join(d, f"graph_{i:04d}.gpickle"), "wb") as f:
                    pickle.dump(g, f)
            ds = TopologicalDataset(directory=d)
            return [ds[i] for i in range(len(ds))], list(ds.FEATURES)
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k.split(".")[0] in ("topological_training", "constants")]:
            del sys.modules[k]


def canonical(graph, data):
    """What the device builder is compared on: node order (conn ids), x, y, the directed edge SET sorted by
    (src, dst) -- the reference's own edge ORDER follows CPython set iteration (to_graph.py:278) and is
    not part of the contract."""
    conn = [int(str(n).split("_")[1]) for n in graph.nodes()]
    ei = data.edge_index
    order = torch.argsort(ei[0] * (len(conn) + 1) + ei[1])
    return {"conn_ids": torch.tensor(conn), "x": data.x.clone(), "y": data.y.clone(),
            "edge_index_sorted": ei[:, order].contiguous()}


def main():
    from gnn_qot_estimation_b200 import synthetic
    gold = {"cases": CASES, "features": FEATURES, "results": []}
    for (S, L, Q, seed, spacing) in CASES:
        samples = synthetic.network_status_samples(S, L, Q, seed=seed, spacing=spacing)
        graphs = reference_graphs(samples)
        datas, names = reference_data_objects(graphs)
        assert names == ["freq", "is_lut", "mod_order", "num_spans", "path_len"], names
        gold["results"].append({"input_checksum": float(np.abs(samples["data"]).sum(dtype=np.float64)),
                                "graphs": [canonical(g, d) for g, d in zip(graphs, datas)]})
        print(f"case {(S, L, Q, seed, spacing)}: nodes {[len(g) for g in graphs]}, "
              f"edges {[g.number_of_edges() for g in graphs]}, self loops {[len(list(__import__('networkx').selfloop_edges(g))) for g in graphs]}")
    torch.save(gold, Path(__file__).parent / "to_graph_lightpath.pt")
    # topological representation: the reference's edge order is part of the contract here
    tcases = [c + (75,) for c in CASES] + [(5, 10, 48, 4, 0.0375, 6)]     # last: 6 endpoints -> duplicate node pairs
    tgold = {"cases": tcases, "features": FEATURES, "results": []}
    for (S, L, Q, seed, spacing, nn) in tcases:
        samples = synthetic.network_status_samples(S, L, Q, seed=seed, spacing=spacing, num_nodes=nn)
        graphs = reference_graphs(samples, representation="topological")
        datas, names = reference_topological_data_objects(graphs)
        assert names == ["freq", "mod_order", "num_spans", "path_len"], names
        tgold["results"].append({"graphs": [{"edge_index": d.edge_index.clone(), "edge_attr": d.edge_attr.clone(),
                                             "y": d.y.clone(), "num_nodes": int(d.num_nodes)} for d in datas]})
        print(f"topological case {(S, L, Q, seed, spacing, nn)}: directed edges {[int(d.edge_index.shape[1]) for d in datas]}")
    torch.save(tgold, Path(__file__).parent / "to_graph_topological.pt")


if __name__ == "__main__":
    main()
