"""Seeded synthetic graph generators for the BASELINE.json configurations.

The reference's dataset (HHI ``.nc`` -> networkx pickles) is not shipped
(SURVEY.md 0.5), so every measurable workload is synthetic, in the exact tensor
layout the reference datasets emit (topological_training/dataset.py:75-123,
lightpath_training/dataset.py:86-123): edges in ``from_networkx`` order (both
directions of every undirected link, grouped by source node ascending).

Generators are vectorised torch code and run on any device (GPU for the 1M-graph
inference workload, CPU for tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .batch import PackedGraphStore

# NSFNET T1 backbone: 14 nodes, 21 undirected links (BASELINE cfg 1/3).
NSFNET_LINKS: List[Tuple[int, int]] = [
    (0, 1), (0, 2), (0, 7), (1, 2), (1, 3), (2, 5), (3, 4), (3, 10), (4, 5), (4, 6),
    (5, 9), (5, 13), (6, 7), (7, 8), (8, 9), (8, 11), (8, 12), (10, 11), (10, 12),
    (11, 13), (12, 13),
]


def directed_from_undirected(num_nodes: int, links: List[Tuple[int, int]]):
    """Directed edge list in ``from_networkx`` order: for each source node ascending,
    its neighbours in adjacency-insertion order.  Also returns, for
    each directed edge, the index of the undirected link it mirrors."""
    adj = [[] for _ in range(num_nodes)]
    for li, (u, v) in enumerate(links):
        adj[u].append((v, li))
        if u != v:
            adj[v].append((u, li))
    src, dst, lid = [], [], []
    for u in range(num_nodes):
        for v, li in adj[u]:
            src.append(u)
            dst.append(v)
            lid.append(li)
    return src, dst, lid


def nsfnet_store(num_graphs: int, seed: int = 0, device="cpu") -> PackedGraphStore:
    """cfg 1/3: `num_graphs` copies of NSFNET (14 nodes, 42 directed edges), edge_attr
    ~ U[0,1) per undirected link mirrored on both directions, y ~ U[0,1) [G,3];
    no node features (node_ids + embeddings)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    src, dst, lid = directed_from_undirected(14, NSFNET_LINKS)
    n_e = len(src)
    link_attr = torch.rand(num_graphs, len(NSFNET_LINKS), 4, generator=gen)
    edge_feat = link_attr[:, torch.tensor(lid)].reshape(num_graphs * n_e, 4).contiguous()
    y = torch.rand(num_graphs, 3, generator=gen)
    node_ptr = torch.arange(num_graphs + 1, dtype=torch.int64) * 14
    edge_ptr = torch.arange(num_graphs + 1, dtype=torch.int64) * n_e
    es = torch.tensor(src, dtype=torch.int32).repeat(num_graphs)
    ed = torch.tensor(dst, dtype=torch.int32).repeat(num_graphs)
    return PackedGraphStore(node_ptr, edge_ptr, es, ed, None, edge_feat, y).to(device)


def lightpath_store(num_graphs: int, seed: int = 1, device="cpu", n_min: int = 8, n_max: int = 56,
                    lut_per_graph: int = 1) -> PackedGraphStore:
    """cfg 2/4: per graph n ~ U{n_min..n_max} lightpath nodes, ~2n random undirected
    interference links (duplicates collapsed like nx.Graph does; mean directed degree
    ~3.8), exactly ``lut_per_graph`` nodes with is_lut == 1.0 (column 1), the other four
    features ~ U[0,1); y ~ U[0,1) [G,3].  Feature order
    [freq, is_lut, mod_order, num_spans, path_len] (lightpath_training/dataset.py:45-48)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    G = num_graphs
    n = torch.randint(n_min, n_max + 1, (G,), generator=gen, device=dev, dtype=torch.int64)
    node_ptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    node_ptr[1:] = torch.cumsum(n, 0)
    N = int(node_ptr[-1])
    m = 2 * n                                                  # undirected samples per graph
    mptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    mptr[1:] = torch.cumsum(m, 0)
    M = int(mptr[-1])
    gid = torch.repeat_interleave(torch.arange(G, device=dev), m)
    ng = n[gid]
    u = (torch.rand(M, generator=gen, device=dev) * ng).long()
    u = torch.minimum(u, ng - 1)
    step = 1 + torch.minimum((torch.rand(M, generator=gen, device=dev) * (ng - 1)).long(), ng - 2)
    v = (u + step) % ng                                        # v != u
    lo, hi = torch.minimum(u, v), torch.maximum(u, v)
    key = (gid * 64 + lo) * 64 + hi                            # n_max <= 63
    key = torch.unique(key)                                    # nx.Graph keeps one edge per pair
    g_u = key // 4096
    lo, hi = (key // 64) % 64, key % 64
    # both directions, grouped by (graph, source) ascending, then by target
    dsrc = torch.cat([lo, hi])
    ddst = torch.cat([hi, lo])
    dg = torch.cat([g_u, g_u])
    order = torch.argsort((dg * 64 + dsrc) * 64 + ddst)
    dsrc, ddst, dg = dsrc[order], ddst[order], dg[order]
    edge_ptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    edge_ptr[1:] = torch.cumsum(torch.bincount(dg, minlength=G), 0)
    x = torch.rand(N, 5, generator=gen, device=dev)
    x[:, 1] = 0.0
    for k in range(lut_per_graph):
        pos = torch.minimum((torch.rand(G, generator=gen, device=dev) * n).long(), n - 1)
        x[node_ptr[:-1] + pos, 1] = 1.0
    y = torch.rand(G, 3, generator=gen, device=dev)
    return PackedGraphStore(node_ptr, edge_ptr, dsrc.to(torch.int32), ddst.to(torch.int32), x, None, y, lut_col=1)


def random_topology_store(num_nodes: int = 10000, num_links: int = 40000, seed: int = 2,
                          device="cpu") -> PackedGraphStore:
    """cfg 5: ONE graph with `num_nodes` nodes and `num_links` random undirected links
    (both directions -> 2*num_links directed edges, grouped by source), edge_attr
    ~ U[0,1) per link mirrored.  No node features (embedding table of num_nodes rows)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    u = torch.randint(0, num_nodes, (num_links,), generator=gen)
    v = (u + 1 + torch.randint(0, num_nodes - 1, (num_links,), generator=gen)) % num_nodes
    attr = torch.rand(num_links, 4, generator=gen)
    src = torch.cat([u, v])
    dst = torch.cat([v, u])
    ea = torch.cat([attr, attr])
    order = torch.argsort(src * num_nodes + dst, stable=True)
    src, dst, ea = src[order], dst[order], ea[order].contiguous()
    node_ptr = torch.tensor([0, num_nodes], dtype=torch.int64)
    edge_ptr = torch.tensor([0, src.numel()], dtype=torch.int64)
    y = torch.rand(1, 3, generator=gen)
    return PackedGraphStore(node_ptr, edge_ptr, src.to(torch.int32), dst.to(torch.int32), None, ea, y).to(device)


# --------------------------------------------------------------------------------------------- #
# raw "network status" samples: the input of to_graph.py (the HHI .nc dataset is not shipped)
# --------------------------------------------------------------------------------------------- #
LP_FEAT = ["conn_id", "src_id", "dst_id", "mod_order", "path_len", "num_spans", "freq", "osnr", "snr", "ber"]
METRICS = ["osnr", "snr", "ber", "class"]


def network_status_samples(num_samples: int, num_links: int = 40, num_freqs: int = 64, seed: int = 0,
                           spacing: float = 0.0375, max_lightpaths: int = 48, super_channel_p: float = 0.1,
                           num_nodes: int = 75):
    """Seeded stand-in for ``dataset["data"]`` of to_graph.py:124-129: a float32 array
    ``[sample, lp_feat, link, freq]`` that is zero on free channels and carries the lightpath's
    feature vector on every (link, frequency) channel it occupies, plus ``target [sample, 4]``,
    the frequency grid ``freqs [num_freqs]`` (float64, 192.2 + k*spacing) and the name lists.
    Every sample holds 8..max_lightpaths lightpaths routed over 1..6 random links with one
    frequency slot each (a few take two adjacent slots: super-channels, which to_graph.py turns
    into self loops); the last lightpath placed is the LUT (osnr = snr = ber = -1).  Endpoints are drawn
    from 1..num_nodes (small values give many lightpaths per node pair: nx.Graph keeps one edge, last
    attributes win, to_graph.py:175-178)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    F = len(LP_FEAT)
    fi = {k: i for i, k in enumerate(LP_FEAT)}
    freqs = 192.2 + spacing * np.arange(num_freqs, dtype=np.float64)
    data = np.zeros((num_samples, F, num_links, num_freqs), dtype=np.float32)
    target = np.zeros((num_samples, len(METRICS)), dtype=np.float64)
    for s in range(num_samples):
        n_lp = int(rng.integers(8, max_lightpaths + 1))
        free = np.ones((num_links, num_freqs), dtype=bool)
        conn_ids = rng.permutation(np.arange(1, 4 * max_lightpaths))[:n_lp]
        placed = 0
        for k in range(n_lp):
            width = 2 if rng.random() < super_channel_p else 1
            links = rng.choice(num_links, size=min(int(rng.integers(1, 7)), num_links), replace=False)
            ok_q = [q for q in range(num_freqs - width + 1) if all(free[l, q:q + width].all() for l in links)]
            if not ok_q:
                continue
            q0 = int(rng.choice(ok_q))
            vec = np.zeros(F, dtype=np.float32)
            vec[fi["conn_id"]] = conn_ids[k]
            vec[fi["src_id"]], vec[fi["dst_id"]] = rng.choice(np.arange(1, num_nodes + 1), 2, replace=False)
            vec[fi["mod_order"]] = rng.choice([4, 8, 16, 32, 64])
            vec[fi["path_len"]] = int(rng.integers(24214, 7834746))
            vec[fi["num_spans"]] = int(rng.integers(1, 107))
            vec[fi["osnr"]], vec[fi["snr"]], vec[fi["ber"]] = rng.uniform(13, 33), rng.uniform(9, 29), rng.uniform(1e-11, 1e-2)
            for l in links:
                for q in range(q0, q0 + width):
                    v = vec.copy()
                    v[fi["freq"]] = freqs[q]
                    data[s, :, l, q] = v
                    free[l, q] = False
            placed += 1
            last = (links, q0, width)
        links, q0, width = last                       # the LUT: metrics unknown (-1), to be predicted
        for l in links:
            for q in range(q0, q0 + width):
                data[s, fi["osnr"], l, q] = data[s, fi["snr"], l, q] = data[s, fi["ber"], l, q] = -1.0
        target[s] = [rng.uniform(12.47, 33.49), rng.uniform(8.96, 29.98), rng.uniform(1.7e-12, 1.98e-2), rng.integers(0, 3)]
    return {"data": data, "target": target, "freqs": freqs, "lp_feat": list(LP_FEAT), "metric": list(METRICS)}
