"""GPU parity, integer side (bit-exact): stable CSR construction, graph / edge offsets
and the device-side collate (csrc/graph_index.cu) against oracle.build_csr_ref /
oracle.collate_ref -- the restatement of PyG's Batch.from_data_list as driven by
topological_training/train.py:93-95 and lightpath_training/train.py:94-96."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand_edges(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)])


@pytest.mark.parametrize("n,e", [(1, 0), (1, 5), (14, 42), (257, 4000), (5000, 20000), (300, 60000)])
@pytest.mark.parametrize("by", [1, 0])
def test_csr_equals_stable_argsort(cuda, n, e, by):
    from gnn_qot_estimation_b200 import ops
    from oracle import build_csr_ref
    ei = _rand_edges(n, e, seed=n + e)
    csr = ops.build_csr(ei.to(cuda), n, by=by, flags=0)
    ref_ei = ei if by == 1 else ei.flip(0)
    rp, nbr, eid = build_csr_ref(ref_ei, n)
    assert int(csr.status.item()) == 0
    assert torch.equal(csr.rowptr.cpu(), rp)
    assert torch.equal(csr.eid.cpu(), eid)
    assert torch.equal(csr.nbr.cpu(), nbr)


def test_csr_hub_rows(cuda):
    """in-degree far above the per-thread insertion-sort limit (rank-sort path)."""
    from gnn_qot_estimation_b200 import ops
    from oracle import build_csr_ref
    n, e = 50, 30000
    ei = _rand_edges(n, e, seed=7)
    ei[1, ::2] = 3          # half of all edges point at node 3
    csr = ops.build_csr(ei.to(cuda), n, by=1, flags=0)
    rp, nbr, eid = build_csr_ref(ei, n)
    assert torch.equal(csr.rowptr.cpu(), rp) and torch.equal(csr.eid.cpu(), eid) and torch.equal(csr.nbr.cpu(), nbr)


@pytest.mark.parametrize("n,e", [(1, 0), (3, 9), (200, 1500), (4096, 30000)])
def test_csr_gat_self_loops(cuda, n, e):
    """flags=3 == remove_self_loops + add_self_loops (GATConv): appended
    loops carry eid = E + node and sort last in their row."""
    from gnn_qot_estimation_b200 import ops
    from oracle import build_csr_ref, gat_edges_ref
    ei = _rand_edges(n, e, seed=3 * n + 1)
    if e:
        ei[1, :: 5] = ei[0, :: 5]          # plant self loops
    csr = ops.build_csr(ei.to(cuda), n, by=1, flags=3)
    gei = gat_edges_ref(ei, n)
    rp, nbr, eid_ref = build_csr_ref(gei, n)
    # map the oracle's positions in the filtered list back to original ids / E+node
    keep = torch.nonzero(ei[0] != ei[1]).flatten()
    orig = torch.cat([keep, e + torch.arange(n)]).to(torch.int32)
    total = int(rp[-1])
    assert torch.equal(csr.rowptr.cpu(), rp)
    assert torch.equal(csr.nbr.cpu()[:total], nbr)
    assert torch.equal(csr.eid.cpu()[:total], orig[eid_ref.long()])


def test_csr_out_of_range_sets_status(cuda):
    from gnn_qot_estimation_b200 import ops
    ei = torch.tensor([[0, 1, 9], [1, 2, 0]])
    csr = ops.build_csr(ei.to(cuda), 3, by=1, flags=0)
    assert int(csr.status.item()) != 0


def test_graph_and_edge_ptr(cuda):
    from gnn_qot_estimation_b200 import ops
    from oracle import graph_ptr_ref
    sizes = [3, 0, 5, 1, 0, 0, 7]           # empty graphs in the middle and at the end
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    gp = ops.graph_ptr(batch.to(cuda), len(sizes))
    assert torch.equal(gp.cpu(), graph_ptr_ref(batch, len(sizes)))
    # edges grouped by graph
    ei = torch.tensor([[0, 1, 2, 3, 4, 9, 10], [1, 2, 0, 4, 3, 10, 9]])
    ep, st = ops.edge_ptr(ei.to(cuda), batch.to(cuda), len(sizes))
    assert int(st.item()) == 0
    assert ep.cpu().tolist() == [0, 3, 3, 5, 5, 5, 5, 7]
    _, st = ops.edge_ptr(ei.flip(1).contiguous().to(cuda), batch.to(cuda), len(sizes))
    assert int(st.item()) != 0


@pytest.mark.parametrize("kind", ["nsfnet", "lightpath"])
@pytest.mark.parametrize("ids", ["range", "shuffled", "repeated"])
def test_collate_equals_pyg_collate(cuda, kind, ids):
    from gnn_qot_estimation_b200 import synthetic
    from oracle import collate_ref
    store = synthetic.nsfnet_store(50, seed=0) if kind == "nsfnet" else synthetic.lightpath_store(50, seed=1)
    if ids == "range":
        sel = range(7, 41)
    elif ids == "shuffled":
        sel = torch.randperm(50, generator=torch.Generator().manual_seed(0))[:33]
    else:
        sel = torch.tensor([4, 4, 9, 0, 4])
    sel_list = list(sel) if isinstance(sel, range) else sel.tolist()
    graphs = []
    for g in sel_list:
        n0, n1 = int(store.node_ptr[g]), int(store.node_ptr[g + 1])
        e0, e1 = int(store.edge_ptr[g]), int(store.edge_ptr[g + 1])
        d = {"num_nodes": n1 - n0,
             "edge_index": torch.stack([store.edge_src[e0:e1], store.edge_dst[e0:e1]]).long(),
             "y": store.y[g:g + 1]}
        if store.node_feat is not None:
            d["x"] = store.node_feat[n0:n1]
        else:
            d["node_ids"] = torch.arange(n1 - n0)
        if store.edge_feat is not None:
            d["edge_attr"] = store.edge_feat[e0:e1]
        graphs.append(d)
    ref = collate_ref(graphs)
    got = store.to(cuda).collate(sel)
    assert got.num_graphs == ref.num_graphs
    for k in ("x", "edge_index", "edge_attr", "batch", "node_ids", "y"):
        a, b = getattr(got, k), getattr(ref, k)
        assert (a is None) == (b is None), k
        if a is not None:
            assert a.dtype == b.dtype and torch.equal(a.cpu(), b), k
    assert torch.equal(got.ptr.cpu(), ref.ptr) and torch.equal(got.edge_ptr.cpu(), ref.edge_ptr)


def test_csr_large_sortedness_property(cuda):
    """Full-size property check (cfg-2 batch: ~131k nodes / ~500k edges): rows are
    grouped by destination, edge ids ascend inside every row, and eid is a permutation."""
    from gnn_qot_estimation_b200 import ops, synthetic
    b = synthetic.lightpath_store(4096, seed=1).host_batch(0, 4096).to(cuda)
    n, e = b.num_nodes, b.num_edges
    csr = ops.build_csr(b.edge_index, n, by=1, flags=0)
    rp, eid, nbr = csr.rowptr.long(), csr.eid.long(), csr.nbr.long()
    assert int(rp[0]) == 0 and int(rp[-1]) == e
    assert torch.equal(torch.sort(eid).values, torch.arange(e, device=cuda))
    row_of_slot = torch.repeat_interleave(torch.arange(n, device=cuda), rp[1:] - rp[:-1])
    assert torch.equal(b.edge_index[1][eid], row_of_slot)
    assert torch.equal(b.edge_index[0][eid], nbr)
    inc = eid[1:] > eid[:-1]
    same_row = row_of_slot[1:] == row_of_slot[:-1]
    assert bool((inc | ~same_row).all())
