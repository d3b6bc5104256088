"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's lightpath graph construction,
to_graph.py::create_lightpath_graph (:187-312), followed by the tensorisation of
lightpath_training/dataset.py::LightpathDataset.__getitem__ (:53-123).

PINNED: unlike the PyG layers, this part of the reference is plain numpy / networkx and RUNS in the
build container (only its ``xarray`` import needs a stand-in), so tests/golden/to_graph_lightpath.pt
holds outputs of the reference's own, unmodified code (tests/golden/make_to_graph_golden.py) and this
restatement is checked against them bit for bit (tests/test_to_graph_cpu.py).

Contract compared on: node ORDER (first appearance of a conn_id in the row-major scan of occupied
(link, freq) channels), x [n,5] float32 in sorted-name order [freq, is_lut, mod_order, num_spans,
path_len] (min-max scaled in float64, then rounded to float32), y [1,3] float32, and the directed edge
SET (both directions of every interaction, a self loop once).  The reference's edge ORDER follows
CPython set iteration (to_graph.py:278) and is not part of the contract: edges are emitted sorted by
(source, target).
"""
import numpy as np

FEATURE_RANGES = {          # constants.py:1-6
    "mod_order": (0.0, 64.0), "path_len": (24214.0, 7834746.0), "num_spans": (1.0, 106.0), "freq": (192.2, 195.8)}
TARGET_RANGES = {           # constants.py:8-12
    "osnr": (12.47, 33.49), "snr": (8.96, 29.98), "ber": (1.70e-12, 1.98e-2)}
NODE_FEATURES = ["freq", "is_lut", "mod_order", "num_spans", "path_len"]   # sorted names, dataset.py:45-48


def lightpath_graph_ref(sample, freqs, lp_feat, freq_threshold=0.05):
    """sample [F, L, Q] (zero = free channel), freqs [Q] float64 -> (conn_ids [n] in node order,
    first_channel [n,2] (link, freq index of the channel the node features come from), is_lut [n],
    undirected edge set as sorted (i, j), i <= j node indices)."""
    fi = {k: i for i, k in enumerate(lp_feat)}
    occupied = np.any(sample != 0, axis=0)                      # to_graph.py:229
    occupied_indices = np.argwhere(occupied)                    # :232, row-major: link-major, freq-minor
    node_of = {}                                                # conn_id -> node index (dict order, :245-268)
    conn_ids, first, is_lut = [], [], []
    per_link = {}                                               # link -> [(freq value, node)]  (:271-275)
    for l, q in occupied_indices:
        vec = sample[:, l, q]
        c = int(vec[fi["conn_id"]])                             # :243, truncation toward zero
        if c not in node_of:
            node_of[c] = len(conn_ids)
            conn_ids.append(c)
            first.append((int(l), int(q)))
            is_lut.append(int(vec[fi["osnr"]] == -1 and vec[fi["snr"]] == -1 and vec[fi["ber"]] == -1))   # :247-251
        per_link.setdefault(int(l), []).append((freqs[q], node_of[c]))
    edges = set()
    for l, ent in per_link.items():                             # :283-310
        if len({n for _, n in ent}) < 2:                        # :285-286: a link used by one lightpath only
            continue
        f = np.array([e[0] for e in ent], dtype=np.float64)
        n = np.array([e[1] for e in ent])
        diff = np.abs(f[:, None] - f[None, :])                  # :296
        ii, jj = np.where((diff < freq_threshold) & (diff > 0))  # :299
        for a, b in zip(n[ii], n[jj]):
            edges.add((int(min(a, b)), int(max(a, b))))
    return (np.array(conn_ids, dtype=np.int64), np.array(first, dtype=np.int64).reshape(-1, 2),
            np.array(is_lut, dtype=np.int64), sorted(edges))


def lightpath_data_ref(sample, target, freqs, lp_feat, metric, freq_threshold=0.05):
    """The tensors LightpathDataset.__getitem__ would return for this sample (numpy): conn_ids, x, y,
    edge_index_sorted [2, E_dir]."""
    fi = {k: i for i, k in enumerate(lp_feat)}
    conn, first, is_lut, und = lightpath_graph_ref(sample, freqs, lp_feat, freq_threshold)
    n = len(conn)
    x = np.zeros((n, len(NODE_FEATURES)), dtype=np.float32)
    for i in range(n):
        l, q = first[i]
        for k, name in enumerate(NODE_FEATURES):
            if name == "is_lut":
                x[i, k] = float(is_lut[i])                      # dataset.py:74-75
            else:
                lo, hi = FEATURE_RANGES[name]
                x[i, k] = (float(sample[fi[name], l, q]) - lo) / (hi - lo)    # :77-80, float64 then float32
    mi = {k: i for i, k in enumerate(metric)}
    y = np.array([[(float(target[mi[k]]) - TARGET_RANGES[k][0]) / (TARGET_RANGES[k][1] - TARGET_RANGES[k][0])
                   for k in ("osnr", "snr", "ber")]], dtype=np.float32)       # :111-121
    d = set()
    for a, b in und:                                            # from_networkx: both directions, a self loop once
        d.add((a, b))
        d.add((b, a))
    d = sorted(d)
    ei = np.array(d, dtype=np.int64).reshape(-1, 2).T if d else np.zeros((2, 0), dtype=np.int64)
    return conn, x, y, np.ascontiguousarray(ei)


TOPO_FEATURES = ["freq", "mod_order", "num_spans", "path_len"]     # sorted names, topological dataset.py:38-41


def topological_data_ref(sample, target, lp_feat, metric, num_nodes=75):
    """to_graph.py::create_topological_graph (:62-184) + TopologicalDataset.__getitem__
    (topological_training/dataset.py:46-123) for one sample -> (edge_index [2,E] int64 in the reference's
    ORDER, edge_attr [E,4] float32, y [3] float32).  Order: lightpaths are added in ascending conn_id
    (np.unique, :147), nx.Graph keeps one edge per node pair -- adjacency position from the FIRST add,
    attributes from the LAST (:175-178); the relabelling copy of dataset.py:57 re-orders the adjacency
    lists (below) and from_networkx lists, for every node ascending, its neighbours in adjacency order."""
    fi = {k: i for i, k in enumerate(lp_feat)}
    occupied = np.any(sample != 0, axis=0)                              # :140
    vecs = sample[:, occupied]                                          # :144, row-major channel order
    conn = vecs[fi["conn_id"]].astype(int)                              # :151
    _, uidx = np.unique(conn, return_index=True)                        # :156: ascending conn_id, first occurrence
    adj = [dict() for _ in range(num_nodes)]                            # adjacency dicts: insertion order
    for k in uidx:
        u, v = int(vecs[fi["src_id"], k]) - 1, int(vecs[fi["dst_id"], k]) - 1     # nodes 1..75 -> 0..74 (dataset.py:57)
        attr = [(float(vecs[fi[name], k]) - FEATURE_RANGES[name][0]) / (FEATURE_RANGES[name][1] - FEATURE_RANGES[name][0])
                for name in TOPO_FEATURES]                              # dataset.py:65-72
        adj[u][v] = attr                                                # nx.Graph.add_edge: both directions share
        adj[v][u] = attr                                                # the (last) attribute dict
    # dataset.py:57 nx.convert_node_labels_to_integers copies the graph by re-adding G.edges() (every node
    # ascending, its not-yet-visited neighbours in adjacency order), which re-orders the adjacency lists:
    # neighbours with a smaller id come first, ascending; then the others in the order G first saw them
    H = [dict() for _ in range(num_nodes)]
    seen = set()
    for u in range(num_nodes):
        for v, attr in adj[u].items():
            if v not in seen:
                H[u][v] = attr
                H[v][u] = attr
        seen.add(u)
    src, dst, ea = [], [], []
    for u in range(num_nodes):                                          # from_networkx: node ascending, adjacency order
        for v, attr in H[u].items():
            src.append(u); dst.append(v); ea.append(attr)
    ei = np.array([src, dst], dtype=np.int64).reshape(2, -1)
    ea = np.array(ea, dtype=np.float32).reshape(-1, len(TOPO_FEATURES))
    mi = {k: i for i, k in enumerate(metric)}
    y = np.array([(float(target[mi[k]]) - TARGET_RANGES[k][0]) / (TARGET_RANGES[k][1] - TARGET_RANGES[k][0])
                  for k in ("osnr", "snr", "ber")], dtype=np.float32)   # dataset.py:109-121
    return ei, ea, y
