"""Device-side graph construction: the B200 counterpart of the reference's ``to_graph.py``.

``create_lightpath_graphs`` turns a batch of raw network-status samples ``[S, lp_feat, link, freq]``
(what ``xr.open_dataset(...)["data"]`` holds, to_graph.py:216-223) straight into the
:class:`PackedGraphStore` the collate kernel consumes -- i.e. it replaces, for all samples at once,
``to_graph.create_lightpath_graph`` (to_graph.py:187-312) -> pickle (store_graphs.py:73-76) ->
``LightpathDataset.__getitem__`` (lightpath_training/dataset.py:53-123).  The arithmetic runs in
``csrc/to_graph.cu`` (qot_lightpath_graph_count / _fill); there is no CPU path.

Node order, node features, labels and the edge SET equal the reference's (tests/test_to_graph_gpu.py
against golden vectors made by the reference's own code); edges are emitted sorted by (source, target)
-- the reference's edge order follows CPython set iteration (to_graph.py:278) and carries no meaning.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .batch import PackedGraphStore
from .ops import check, ptr, stream

# constants.py:1-12 of the reference
FEATURE_RANGES = {"mod_order": (0.0, 64.0), "path_len": (24214.0, 7834746.0), "num_spans": (1.0, 106.0),
                  "freq": (192.2, 195.8)}
TARGET_RANGES = {"osnr": (12.47, 33.49), "snr": (8.96, 29.98), "ber": (1.70e-12, 1.98e-2)}
NODE_FEATURES = ["freq", "is_lut", "mod_order", "num_spans", "path_len"]     # sorted names, dataset.py:45-48


def _cfg(lp_feat: Sequence[str], metric: Sequence[str], F: int, L: int, Q: int, T: int, freq_threshold: float):
    fi = {k: i for i, k in enumerate(lp_feat)}
    mi = {k: i for i, k in enumerate(metric)}
    for k in ("conn_id", "osnr", "snr", "ber", "freq", "mod_order", "num_spans", "path_len"):
        if k not in fi:
            raise KeyError(f"lp_feat has no '{k}' entry (to_graph.py:27-33 needs it)")
    cfg = _lib.QotLpGraphCfg()
    cfg.F, cfg.L, cfg.Q, cfg.T = F, L, Q, T
    cfg.i_conn, cfg.i_osnr, cfg.i_snr, cfg.i_ber = fi["conn_id"], fi["osnr"], fi["snr"], fi["ber"]
    for q, name in enumerate(("freq", "mod_order", "num_spans", "path_len")):
        cfg.i_feat[q] = fi[name]
        cfg.feat_lo[q], cfg.feat_hi[q] = FEATURE_RANGES[name]
    for q, name in enumerate(("osnr", "snr", "ber")):
        cfg.i_tgt[q] = mi[name]
        cfg.tgt_lo[q], cfg.tgt_hi[q] = TARGET_RANGES[name]
    cfg.freq_threshold = float(freq_threshold)
    return cfg


@_lib.on_tensor_device
def create_lightpath_graphs(data: torch.Tensor, target: torch.Tensor, freqs: torch.Tensor, lp_feat: Sequence[str],
                            metric: Sequence[str], freq_threshold: float = 0.05,
                            return_conn_ids: bool = False):
    """data [S, F, L, Q] float32 (CUDA), target [S, T] float64, freqs [Q] float64 -> PackedGraphStore with
    node_feat [N,5] (``NODE_FEATURES`` order, min-max scaled), y [S,3], lut_col = 1.  The sample tensor is read
    once (scan launch -> per-sample records), then packed; one host sync (the per-sample counts are scanned on
    the device, their totals size the output)."""
    if not data.is_cuda:
        raise RuntimeError("create_lightpath_graphs needs CUDA tensors (gnn_qot_estimation_b200 has no CPU path)")
    dev = data.device
    data = data.to(torch.float32).contiguous()
    target = target.to(dev, torch.float64).contiguous()
    freqs = freqs.to(dev, torch.float64).contiguous()
    S, F, L, Q = (int(v) for v in data.shape)
    if freqs.numel() != Q or target.shape[0] != S:
        raise ValueError("freqs must have data.shape[3] entries and target data.shape[0] rows")
    cfg = _cfg(lp_feat, metric, F, L, Q, int(target.shape[1]), freq_threshold)
    lib = _lib.lib()
    counts = torch.zeros(S, 2, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    y = torch.empty(S, 3, dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.qot_lightpath_graph_scratch_bytes(S), dtype=torch.uint8, device=dev)
    check(lib.qot_lightpath_graph_count(ptr(data), ptr(freqs), ptr(target), S, C.byref(cfg), ptr(counts), ptr(y),
                                        ptr(scratch), scratch.numel(), ptr(status), stream()),
          "qot_lightpath_graph_count")
    ptrs = torch.zeros(2, S + 1, dtype=torch.int64, device=dev)
    ptrs[:, 1:] = torch.cumsum(counts.to(torch.int64).t(), dim=1)
    n_tot, e_tot, st = (int(v) for v in torch.stack([ptrs[0, -1], ptrs[1, -1], status[0].to(torch.int64)]).tolist())
    if st & 1:
        raise RuntimeError("create_lightpath_graphs: a sample exceeds the per-block capacity "
                           "(QOT_TG_MAX_CHANNELS occupied channels / QOT_TG_MAX_NODES lightpaths / QOT_TG_MAX_LINKS links)")
    node_feat = torch.empty(max(n_tot, 1), 5, dtype=torch.float32, device=dev)[:n_tot]
    conn_ids = torch.empty(max(n_tot, 1), dtype=torch.int64, device=dev)[:n_tot]
    edge_src = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
    edge_dst = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
    node_ptr, edge_ptr = ptrs[0].contiguous(), ptrs[1].contiguous()
    check(lib.qot_lightpath_graph_fill(ptr(scratch), S, ptr(node_ptr), ptr(edge_ptr), ptr(node_feat), ptr(conn_ids),
                                       ptr(edge_src), ptr(edge_dst), ptr(status), stream()), "qot_lightpath_graph_fill")
    store = PackedGraphStore(node_ptr, edge_ptr, edge_src, edge_dst, node_feat, None, y, lut_col=1)
    return (store, conn_ids) if return_conn_ids else store


@_lib.on_tensor_device
def create_topological_graphs(data: torch.Tensor, target: torch.Tensor, lp_feat: Sequence[str], metric: Sequence[str],
                              num_nodes: int = 75) -> PackedGraphStore:
    """The topological representation for all samples at once: ``to_graph.create_topological_graph``
    (to_graph.py:62-184) -> pickle -> ``TopologicalDataset.__getitem__`` (topological_training/dataset.py:46-123).
    data [S, F, L, Q] float32 (CUDA), target [S, T] float64 -> PackedGraphStore with ``num_nodes`` nodes per
    graph (no node features: the model embeds ``node_ids``), edges in the reference's order, edge_feat [E,4] in
    sorted-name order [freq, mod_order, num_spans, path_len] (min-max scaled), y [S,3]."""
    if not data.is_cuda:
        raise RuntimeError("create_topological_graphs needs CUDA tensors (gnn_qot_estimation_b200 has no CPU path)")
    dev = data.device
    data = data.to(torch.float32).contiguous()
    target = target.to(dev, torch.float64).contiguous()
    S, F, L, Q = (int(v) for v in data.shape)
    if target.shape[0] != S:
        raise ValueError("target must have data.shape[0] rows")
    fi = {k: i for i, k in enumerate(lp_feat)}
    for k in ("src_id", "dst_id"):
        if k not in fi:
            raise KeyError(f"lp_feat has no '{k}' entry (to_graph.py:152-153 needs it)")
    cfg = _cfg(lp_feat, metric, F, L, Q, int(target.shape[1]), 0.05)
    lib = _lib.lib()
    counts = torch.zeros(S, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    y = torch.empty(S, 3, dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.qot_topological_graph_scratch_bytes(S), dtype=torch.uint8, device=dev)
    check(lib.qot_topological_graph_count(ptr(data), ptr(target), S, C.byref(cfg), int(num_nodes), fi["src_id"],
                                          fi["dst_id"], ptr(counts), ptr(y), ptr(scratch), scratch.numel(),
                                          ptr(status), stream()), "qot_topological_graph_count")
    edge_ptr = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    edge_ptr[1:] = torch.cumsum(counts.to(torch.int64), dim=0)
    e_tot, st = (int(v) for v in torch.stack([edge_ptr[-1], status[0].to(torch.int64)]).tolist())
    if st & 1:
        raise RuntimeError("create_topological_graphs: a sample exceeds the per-block capacity or names a network "
                           f"node outside 1..{num_nodes}")
    edge_src = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
    edge_dst = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
    edge_feat = torch.empty(max(e_tot, 1), 4, dtype=torch.float32, device=dev)[:e_tot]
    check(lib.qot_topological_graph_fill(ptr(scratch), S, ptr(edge_ptr), ptr(edge_src), ptr(edge_dst), ptr(edge_feat),
                                         ptr(status), stream()), "qot_topological_graph_fill")
    node_ptr = torch.arange(S + 1, dtype=torch.int64, device=dev) * int(num_nodes)
    return PackedGraphStore(node_ptr, edge_ptr, edge_src, edge_dst, None, edge_feat, y)
