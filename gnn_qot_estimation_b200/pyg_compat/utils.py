"""``torch_geometric.utils.from_networkx`` as the reference uses it (*/dataset.py:75|86)."""
from __future__ import annotations

import torch

from .data import Data


def from_networkx(G) -> Data:
    """Undirected graphs become directed (both directions of every edge; a self loop once); nodes are
    relabelled 0..n-1 in ``G.nodes()`` order; ``edge_index`` follows ``G.edges()`` -- grouped by source
    in ascending node order, neighbours in adjacency-insertion order.  Node / edge /
    graph attributes that convert to tensors are attached under their names (the reference overwrites
    ``x`` / ``edge_attr`` / ``y`` itself right after)."""
    import networkx as nx
    G = G.to_directed() if not nx.is_directed(G) else G
    mapping = {node: i for i, node in enumerate(G.nodes())}
    edges = [(mapping[u], mapping[v]) for u, v in G.edges()]
    data = Data()
    data.edge_index = torch.tensor(edges, dtype=torch.int64).t().contiguous().view(2, -1)
    data.num_nodes = G.number_of_nodes()

    def attach(name, values):
        try:
            t = torch.tensor(values)
        except (ValueError, TypeError, RuntimeError):
            return
        setattr(data, name, t)

    node_attrs = [a for _, a in G.nodes(data=True)]
    if node_attrs and node_attrs[0]:
        for key in node_attrs[0]:
            if all(key in a for a in node_attrs):
                attach(str(key), [a[key] for a in node_attrs])
    edge_attrs = [a for _, _, a in G.edges(data=True)]
    if edge_attrs and edge_attrs[0]:
        for key in edge_attrs[0]:
            if all(key in a for a in edge_attrs):
                name = str(key)
                attach(f"edge_{name}" if name in data else name, [a[key] for a in edge_attrs])
    for key, value in G.graph.items():
        if str(key) not in data:
            attach(str(key), value)
    return data
