"""GPU parity, SURVEY 8(f)4: csrc/to_graph.cu through gnn_qot_estimation_b200.to_graph.create_lightpath_graphs
against (a) golden vectors made by the REFERENCE's own to_graph.py + LightpathDataset and (b) the numpy
oracle on fresh seeds.  Bit-exact: node order, x, y, edge set."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _build(samples, dev):
    from gnn_qot_estimation_b200.to_graph import create_lightpath_graphs
    return create_lightpath_graphs(torch.from_numpy(samples["data"]).to(dev), torch.from_numpy(samples["target"]),
                                   torch.from_numpy(samples["freqs"]), samples["lp_feat"], samples["metric"],
                                   return_conn_ids=True)


def _graph(store, conn, i):
    n0, n1 = int(store.node_ptr[i]), int(store.node_ptr[i + 1])
    e0, e1 = int(store.edge_ptr[i]), int(store.edge_ptr[i + 1])
    ei = torch.stack([store.edge_src[e0:e1], store.edge_dst[e0:e1]]).to(torch.int64).cpu()
    return conn[n0:n1].cpu(), store.node_feat[n0:n1].cpu(), store.y[i:i + 1].cpu(), ei


def test_matches_reference_golden_vectors(cuda):
    from gnn_qot_estimation_b200 import synthetic
    gold = load_golden("to_graph_lightpath.pt")
    for (S, L, Q, seed, spacing), res in zip(gold["cases"], gold["results"]):
        samples = synthetic.network_status_samples(S, L, Q, seed=seed, spacing=spacing)
        store, conn = _build(samples, cuda)
        assert store.num_graphs == S
        for i, g in enumerate(res["graphs"]):
            c, x, y, ei = _graph(store, conn, i)
            assert torch.equal(c, g["conn_ids"])                           # node order = first appearance
            assert torch.equal(x, g["x"]) and torch.equal(y, g["y"])       # bit-exact (fp64 scaling, then fp32)
            assert torch.equal(ei, g["edge_index_sorted"])                 # edge set, incl. self loops and the 0.05 boundary


@pytest.mark.parametrize("S,L,Q,seed,spacing", [(64, 16, 80, 5, 0.0375), (7, 3, 17, 6, 0.05), (5, 120, 96, 7, 0.01)])
def test_matches_oracle_on_fresh_seeds(cuda, S, L, Q, seed, spacing):
    from gnn_qot_estimation_b200 import synthetic
    from oracle import lightpath_data_ref
    samples = synthetic.network_status_samples(S, L, Q, seed=seed, spacing=spacing)
    samples["data"][S // 2] = 0.0                                          # an empty sample: zero nodes, zero edges
    store, conn = _build(samples, cuda)
    for i in range(S):
        ec, ex, ey, eei = lightpath_data_ref(samples["data"][i], samples["target"][i], samples["freqs"],
                                             samples["lp_feat"], samples["metric"])
        c, x, y, ei = _graph(store, conn, i)
        assert np.array_equal(c.numpy(), ec) and np.array_equal(x.numpy(), ex) and np.array_equal(y.numpy(), ey)
        assert np.array_equal(ei.numpy(), eei)


def test_built_store_feeds_the_model(cuda):
    """raw samples -> device graph construction -> device collate -> fused eval kernel == oracle model on the
    oracle-built graphs."""
    from gnn_qot_estimation_b200 import Batch, LightpathGNN, synthetic
    from oracle import LightpathGNNOracle, lightpath_data_ref
    samples = synthetic.network_status_samples(40, 12, 64, seed=9)
    store, _ = _build(samples, cuda)
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    with torch.no_grad():
        out, lb = m(store.collate(range(40)))
    xs, eis, bts, off = [], [], [], 0
    for i in range(40):
        _, x, _, ei = lightpath_data_ref(samples["data"][i], samples["target"][i], samples["freqs"], samples["lp_feat"], samples["metric"])
        xs.append(torch.from_numpy(x)); eis.append(torch.from_numpy(ei) + off); bts.append(torch.full((x.shape[0],), i)); off += x.shape[0]
    ob = Batch(x=torch.cat(xs).double(), edge_index=torch.cat(eis, 1), batch=torch.cat(bts), num_graphs=40)
    om = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0).double()
    om.load_state_dict(sd, strict=True)
    with torch.no_grad():
        eo, el = om.eval()(ob)
    assert torch.equal(lb.cpu(), el)
    assert float((out.cpu().double() - eo).abs().max() / eo.abs().max()) <= 1e-5


def test_capacity_overflow_is_reported(cuda):
    from gnn_qot_estimation_b200.to_graph import create_lightpath_graphs
    from gnn_qot_estimation_b200 import synthetic
    F = len(synthetic.LP_FEAT)
    data = torch.zeros(1, F, 100, 80)
    data[0, 0] = torch.arange(8000, dtype=torch.float32).view(100, 80) + 1     # 8000 occupied channels, all distinct ids
    with pytest.raises(RuntimeError, match="capacity"):
        create_lightpath_graphs(data.to(cuda), torch.zeros(1, 4, dtype=torch.float64), torch.linspace(192.2, 195.8, 80, dtype=torch.float64),
                                synthetic.LP_FEAT, synthetic.METRICS)


def test_topological_matches_reference_golden_vectors(cuda):
    """create_topological_graph + TopologicalDataset: edge ORDER, attributes (last lightpath of a node pair
    wins), labels -- bit for bit against the reference's own outputs."""
    from gnn_qot_estimation_b200 import synthetic
    from gnn_qot_estimation_b200.to_graph import create_topological_graphs
    gold = load_golden("to_graph_topological.pt")
    for (S, L, Q, seed, spacing, nn), res in zip(gold["cases"], gold["results"]):
        samples = synthetic.network_status_samples(S, L, Q, seed=seed, spacing=spacing, num_nodes=nn)
        store = create_topological_graphs(torch.from_numpy(samples["data"]).to(cuda), torch.from_numpy(samples["target"]),
                                          samples["lp_feat"], samples["metric"])
        assert store.num_graphs == S and int(store.node_ptr[-1]) == 75 * S
        for i, g in enumerate(res["graphs"]):
            e0, e1 = int(store.edge_ptr[i]), int(store.edge_ptr[i + 1])
            ei = torch.stack([store.edge_src[e0:e1], store.edge_dst[e0:e1]]).to(torch.int64).cpu()
            assert torch.equal(ei, g["edge_index"])
            assert torch.equal(store.edge_feat[e0:e1].cpu(), g["edge_attr"]) and torch.equal(store.y[i].cpu(), g["y"])


@pytest.mark.parametrize("S,L,Q,seed,nn", [(48, 14, 64, 21, 75), (9, 6, 40, 22, 4), (6, 50, 80, 23, 12)])
def test_topological_matches_oracle_on_fresh_seeds(cuda, S, L, Q, seed, nn):
    from gnn_qot_estimation_b200 import synthetic
    from gnn_qot_estimation_b200.to_graph import create_topological_graphs
    from oracle import topological_data_ref
    samples = synthetic.network_status_samples(S, L, Q, seed=seed, num_nodes=nn)
    samples["data"][S // 3] = 0.0
    store = create_topological_graphs(torch.from_numpy(samples["data"]).to(cuda), torch.from_numpy(samples["target"]),
                                      samples["lp_feat"], samples["metric"])
    for i in range(S):
        eei, eea, ey = topological_data_ref(samples["data"][i], samples["target"][i], samples["lp_feat"], samples["metric"])
        e0, e1 = int(store.edge_ptr[i]), int(store.edge_ptr[i + 1])
        ei = torch.stack([store.edge_src[e0:e1], store.edge_dst[e0:e1]]).to(torch.int64).cpu().numpy()
        assert np.array_equal(ei, eei) and np.array_equal(store.edge_feat[e0:e1].cpu().numpy(), eea)
        assert np.array_equal(store.y[i].cpu().numpy(), ey)


def test_topological_store_feeds_the_model(cuda):
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    from gnn_qot_estimation_b200.to_graph import create_topological_graphs
    samples = synthetic.network_status_samples(16, 12, 64, seed=31)
    store = create_topological_graphs(torch.from_numpy(samples["data"]).to(cuda), torch.from_numpy(samples["target"]),
                                      samples["lp_feat"], samples["metric"])
    sd = load_golden("ckpt_topological_model_0.pt")["model_state_dict"]
    m = TopologicalGNN(75, 16, 3, edge_dim=4, dropout_p=0.0)
    m.load_state_dict(sd, strict=True)
    out = m.to(cuda).eval()(store.collate(range(16)))
    assert out.shape == (16, 3) and bool(torch.isfinite(out).all())


def test_sample_order_equivariance_and_batch_invariance(cuda):
    """512 samples [10, 60, 80]: permuting the sample axis permutes the graphs (bit for bit, both
    representations); one call on all samples == one call per chunk, concatenated."""
    from gnn_qot_estimation_b200 import synthetic
    from gnn_qot_estimation_b200.to_graph import create_lightpath_graphs, create_topological_graphs
    s = synthetic.network_status_samples(64, 60, 80, seed=41)
    data = torch.from_numpy(s["data"]).to(cuda).repeat(8, 1, 1, 1)
    data[64:] += 0.0
    # make the repeats differ: shift conn ids per copy (keeps structure, changes labels)
    ci = s["lp_feat"].index("conn_id")
    for r in range(1, 8):
        blk = data[64 * r:64 * (r + 1), ci]
        blk[blk != 0] += 1000.0 * r
    tgt = torch.from_numpy(s["target"]).repeat(8, 1)
    fr = torch.from_numpy(s["freqs"])
    S = data.shape[0]
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(1))

    def graphs(store, with_x):
        out = []
        for i in range(store.num_graphs):
            n0, n1, e0, e1 = int(store.node_ptr[i]), int(store.node_ptr[i + 1]), int(store.edge_ptr[i]), int(store.edge_ptr[i + 1])
            out.append((store.node_feat[n0:n1] if with_x else store.edge_feat[e0:e1], store.edge_src[e0:e1], store.edge_dst[e0:e1], store.y[i]))
        return out

    for build, with_x in ((lambda d, t: create_lightpath_graphs(d, t, fr, s["lp_feat"], s["metric"]), True),
                          (lambda d, t: create_topological_graphs(d, t, s["lp_feat"], s["metric"]), False)):
        base = graphs(build(data, tgt), with_x)
        pg = graphs(build(data[perm.to(cuda)].contiguous(), tgt[perm]), with_x)
        for j, i in enumerate(perm.tolist()):
            assert all(torch.equal(a, b) for a, b in zip(pg[j], base[i]))
        parts = graphs(build(data[:200].contiguous(), tgt[:200]), with_x) + graphs(build(data[200:].contiguous(), tgt[200:]), with_x)
        assert len(parts) == S and all(torch.equal(a, b) for g, h in zip(parts, base) for a, b in zip(g, h))


def test_corner_cases_match_oracle(cuda):
    """The hand-built samples of tests/tg_cases.py (checked on the CPU against the reference's own code): single
    lightpath, the strict 0.05 boundary, a skipped single-lightpath link, a self loop, float conn ids, an empty
    sample -- both representations, bit for bit."""
    from tg_cases import corner_case_samples
    from gnn_qot_estimation_b200.to_graph import create_topological_graphs
    from oracle import lightpath_data_ref, topological_data_ref
    s = corner_case_samples()
    store, conn = _build(s, cuda)
    tstore = create_topological_graphs(torch.from_numpy(s["data"]).to(cuda), torch.from_numpy(s["target"]), s["lp_feat"], s["metric"])
    for i in range(s["data"].shape[0]):
        ec, ex, ey, eei = lightpath_data_ref(s["data"][i], s["target"][i], s["freqs"], s["lp_feat"], s["metric"])
        c, x, y, ei = _graph(store, conn, i)
        assert np.array_equal(c.numpy(), ec) and np.array_equal(x.numpy(), ex) and np.array_equal(ei.numpy(), eei), i
        assert np.array_equal(y.numpy(), ey)
        tei, tea, ty = topological_data_ref(s["data"][i], s["target"][i], s["lp_feat"], s["metric"])
        e0, e1 = int(tstore.edge_ptr[i]), int(tstore.edge_ptr[i + 1])
        gei = torch.stack([tstore.edge_src[e0:e1], tstore.edge_dst[e0:e1]]).to(torch.int64).cpu().numpy()
        assert np.array_equal(gei, tei) and np.array_equal(tstore.edge_feat[e0:e1].cpu().numpy(), tea.reshape(-1, 4)), i
