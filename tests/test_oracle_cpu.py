"""CPU tests that pin the oracle (oracle/qot_oracle.py) -- the checker every GPU parity test
trusts.  The reference ships no tests or golden vectors for this path and its arithmetic
lives in torch_geometric (absent), so the oracle is anchored on:
  (i)   strict load of the three shipped checkpoints (names / shapes / dtypes),
  (ii)  an INDEPENDENT naive restatement (pure-Python loops over edges, no scatter ops) of the
        published PyG layer semantics on small cases + literal
        hand-computed answers,
  (iii) fp64 gradcheck, direct == factorised NNConv,
  (iv)  the committed golden vectors (tests/golden/*.pt) being reproduced bit-for-bit-stable
        by the oracle on this machine (within fp32 round-off).
"""
import math

import numpy as np
import pytest
import torch

import oracle
from conftest import batch_from_dict, grad_errs, load_golden, rel_err

torch.set_num_threads(4)


# ----------------------------------------------------------------------------- naive loops
def naive_transformer_conv(x, ei, ea, Wq, bq, Wk, bk, Wv, bv, We, Ws, bs):
    N, C = x.shape[0], Wq.shape[0]
    lin = lambda W, b, v: [sum(W[o][i] * v[i] for i in range(len(v))) + (b[o] if b is not None else 0.0)
                           for o in range(len(W))]
    X = x.tolist()
    q = [lin(Wq.tolist(), bq.tolist(), X[n]) for n in range(N)]
    k = [lin(Wk.tolist(), bk.tolist(), X[n]) for n in range(N)]
    v = [lin(Wv.tolist(), bv.tolist(), X[n]) for n in range(N)]
    s = [lin(Ws.tolist(), bs.tolist(), X[n]) for n in range(N)]
    e = [lin(We.tolist(), None, a) for a in ea.tolist()]
    out = []
    for i in range(N):
        ins = [t for t in range(ei.shape[1]) if int(ei[1, t]) == i]
        logits = [sum(q[i][c] * (k[int(ei[0, t])][c] + e[t][c]) for c in range(C)) / math.sqrt(C) for t in ins]
        row = list(s[i])
        if ins:
            m = max(logits)
            z = [math.exp(l - m) for l in logits]
            den = sum(z) + 1e-16
            for t, zz in zip(ins, z):
                j = int(ei[0, t])
                for c in range(C):
                    row[c] += zz / den * (v[j][c] + e[t][c])
        out.append(row)
    return torch.tensor(out, dtype=torch.float64)


def naive_nnconv_mean(x, ei, ea, W1, b1, W2, b2, Wroot, bias):
    N, H = x.shape
    X, W1l, b1l, W2l, b2l = x.tolist(), W1.tolist(), b1.tolist(), W2.tolist(), b2.tolist()
    out = []
    for i in range(N):
        ins = [t for t in range(ei.shape[1]) if int(ei[1, t]) == i]
        acc = [0.0] * H
        for t in ins:
            a = ea[t].tolist()
            h = [max(0.0, sum(W1l[k][d] * a[d] for d in range(4)) + b1l[k]) for k in range(len(W1l))]
            j = int(ei[0, t])
            for o in range(H):
                for r in range(H):
                    w = sum(W2l[r * H + o][k] * h[k] for k in range(len(h))) + b2l[r * H + o]
                    acc[o] += X[j][r] * w
        cnt = max(len(ins), 1)
        out.append([acc[o] / cnt + sum(Wroot[o][r].item() * X[i][r] for r in range(H)) + bias[o].item()
                    for o in range(H)])
    return torch.tensor(out, dtype=torch.float64)


def naive_gat(x, ei, W, att_src, att_dst, bias):
    N = x.shape[0]
    heads, C = att_src.shape[-2], att_src.shape[-1]
    X = x.tolist()
    Wl = W.tolist()
    xp = [[sum(Wl[o][f] * X[n][f] for f in range(len(X[n]))) for o in range(heads * C)] for n in range(N)]
    a_s, a_d = att_src.view(heads, C).tolist(), att_dst.view(heads, C).tolist()
    edges = [(int(ei[0, t]), int(ei[1, t])) for t in range(ei.shape[1]) if int(ei[0, t]) != int(ei[1, t])]
    edges += [(n, n) for n in range(N)]
    out = []
    for i in range(N):
        row = bias.tolist()
        ins = [j for (j, d) in edges if d == i]
        for h in range(heads):
            sd = sum(xp[i][h * C + c] * a_d[h][c] for c in range(C))
            logits = []
            for j in ins:
                a = sum(xp[j][h * C + c] * a_s[h][c] for c in range(C)) + sd
                logits.append(a if a > 0 else 0.2 * a)
            m = max(logits)
            z = [math.exp(l - m) for l in logits]
            den = sum(z) + 1e-16
            for j, zz in zip(ins, z):
                for c in range(C):
                    row[h * C + c] += zz / den * xp[j][h * C + c]
        out.append(row)
    return torch.tensor(out, dtype=torch.float64)


def _rand_graph(n, e, seed, self_loops=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)])
    if not self_loops:
        ei = ei[:, ei[0] != ei[1]]
    return ei, g


# ----------------------------------------------------------------------------- (ii) KATs
def test_transformer_conv_hand_case():
    """3 nodes, C=2, identity q/k/v, zero skip; edges 0->2 (attr e=[.5,0]) and 2->2 (self loop
    kept by TransformerConv).  q2=[1,1]; k02=[1.5,0], k22=[1,1] -> logits 1.5/sqrt2, 2/sqrt2."""
    x = torch.tensor([[1., 0.], [0., 1.], [1., 1.]], dtype=torch.float64)
    I, z = torch.eye(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    We = torch.tensor([[1., 0, 0, 0], [0, 1., 0, 0]], dtype=torch.float64)
    ei = torch.tensor([[0, 2], [2, 2]])
    ea = torch.tensor([[.5, 0, 0, 0], [0, 0, 0, 0]], dtype=torch.float64)
    out = oracle.transformer_conv_ref(x, ei, ea, I, z, I, z, I, z, We, torch.zeros(2, 2, dtype=torch.float64), z)
    a0, a2 = 1.5 / math.sqrt(2), 2 / math.sqrt(2)
    w0 = math.exp(a0 - a2) / (math.exp(a0 - a2) + 1.0)
    w2 = 1.0 - w0
    exp = torch.tensor([[0, 0], [0, 0], [w0 * 1.5 + w2 * 1.0, w2 * 1.0]], dtype=torch.float64)
    assert torch.allclose(out, exp, atol=1e-14)


def test_nnconv_hand_case():
    """H=2.  Edge MLP: W1 picks a[0] into h[0] (others 0), b1=0; W2 maps h[0] to the flat [in,out]
    weight [[h,0],[0,2h]], b2=0.  Edges 0->1 (a0=1), 2->1 (a0=3) -> mean of x_j W_e; node 0, 2
    have no in-edges -> root + bias only."""
    x = torch.tensor([[1., 2.], [5., 7.], [3., 4.]], dtype=torch.float64)
    W1 = torch.zeros(8, 4, dtype=torch.float64); W1[0, 0] = 1
    b1 = torch.zeros(8, dtype=torch.float64)
    W2 = torch.zeros(4, 8, dtype=torch.float64); W2[0, 0] = 1; W2[3, 0] = 2
    b2 = torch.zeros(4, dtype=torch.float64)
    Wroot = torch.tensor([[1., 0.], [0., -1.]], dtype=torch.float64)
    bias = torch.tensor([.25, .5], dtype=torch.float64)
    ei = torch.tensor([[0, 2], [1, 1]])
    ea = torch.tensor([[1., 0, 0, 0], [3., 0, 0, 0]], dtype=torch.float64)
    out = oracle.nnconv_mean_ref(x, ei, ea, W1, b1, W2, b2, Wroot, bias)
    # msgs: [1*1, 2*2] = [1,4]; [3*3, 4*6] = [9,24]; mean = [5,14]
    exp = torch.tensor([[1.25, -1.5], [5 + 5 + .25, 14 - 7 + .5], [3.25, -3.5]], dtype=torch.float64)
    assert torch.allclose(out, exp, atol=1e-14)
    out_f = oracle.nnconv_mean_factorised_ref(x, ei, ea, W1, b1, W2, b2, Wroot, bias)
    assert torch.allclose(out_f, exp, atol=1e-14)


def test_gat_hand_case():
    """1 head, C=1, in=1, W=[[2]], att_src=att_dst=1: x'=2x; logits leaky(2x_j+2x_i).
    Graph: 0->1 plus an input self loop 1->1 that GAT drops and re-adds once."""
    x = torch.tensor([[1.], [-2.]], dtype=torch.float64)
    W = torch.tensor([[2.]], dtype=torch.float64)
    att = torch.ones(1, 1, 1, dtype=torch.float64)
    ei = torch.tensor([[0, 1], [1, 1]])
    out = oracle.gat_conv_ref(x, ei, W, att, att, torch.tensor([.5], dtype=torch.float64))
    # node 0: only its self loop -> x'_0 = 2 ; node 1: in-edges {0, self}: logits leaky(2-4)=-.4, leaky(-8)=-1.6
    z0, z1 = math.exp(-.4), math.exp(-1.6)
    exp = torch.tensor([[2 + .5], [(z0 * 2 + z1 * -4) / (z0 + z1) + .5]], dtype=torch.float64)
    assert torch.allclose(out, exp, atol=1e-14)


def test_pool_and_lut_hand_case():
    x = torch.tensor([[1., 2.], [3., 4.], [10., 20.]])
    batch = torch.tensor([0, 0, 2])                          # graph 1 is empty -> zeros (count clamp)
    assert torch.equal(oracle.global_mean_pool_ref(x, batch), torch.tensor([[2., 3.], [0., 0.], [10., 20.]]))
    feat = torch.tensor([[0., 1.], [0., 0.], [0., 1.]])
    h, lb = oracle.lut_select_ref(feat, x, batch, 1)
    assert torch.equal(h, x[[0, 2]]) and torch.equal(lb, torch.tensor([0, 2]))
    with pytest.raises(ValueError, match="No LUT node found in the batch."):
        oracle.lut_select_ref(torch.zeros(3, 2), x, batch, 1)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_layers_match_naive_loops(seed):
    n, e, H = 7, 19, 4
    ei, g = _rand_graph(n, e, seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x, ea = r(n, H), torch.rand(ei.shape[1], 4, generator=g, dtype=torch.float64)
    args = (r(H, H), r(H), r(H, H), r(H), r(H, H), r(H), r(H, 4), r(H, H), r(H))
    assert torch.allclose(oracle.transformer_conv_ref(x, ei, ea, *args), naive_transformer_conv(x, ei, ea, *args), atol=1e-12)
    nargs = (r(8, 4), r(8), r(H * H, 8), r(H * H), r(H, H), r(H))
    exp = naive_nnconv_mean(x, ei, ea, *nargs)
    assert torch.allclose(oracle.nnconv_mean_ref(x, ei, ea, *nargs), exp, atol=1e-12)
    assert torch.allclose(oracle.nnconv_mean_factorised_ref(x, ei, ea, *nargs, chunk=5), exp, atol=1e-12)
    x5 = r(n, 5)
    gargs = (r(8, 5), r(1, 4, 2), r(1, 4, 2), r(8))
    assert torch.allclose(oracle.gat_conv_ref(x5, ei, *gargs), naive_gat(x5, ei, *gargs), atol=1e-12)


def test_segment_softmax_and_scatter_conventions():
    src = torch.tensor([1., 2., 3., -1.], dtype=torch.float64)
    idx = torch.tensor([0, 0, 2, 2])
    sm = oracle.segment_softmax(src, idx, 4)
    e = math.e
    assert torch.allclose(sm, torch.tensor([1 / (1 + e), e / (1 + e), 1 / (1 + e ** -4), e ** -4 / (1 + e ** -4)],
                                           dtype=torch.float64), atol=1e-15)
    assert torch.equal(oracle.scatter_sum(src, idx, 4), torch.tensor([3., 0., 2., 0.], dtype=torch.float64))
    assert torch.equal(oracle.scatter_amax(src, idx, 4), torch.tensor([2., 0., 3., 0.], dtype=torch.float64))


# ----------------------------------------------------------------------------- (iii) identities
def test_gradcheck_layers():
    n, H = 5, 2
    ei, g = _rand_graph(n, 11, 3)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).requires_grad_(True)
    ea = torch.rand(ei.shape[1], 4, generator=g, dtype=torch.float64)
    targs = (r(n, H), r(H, H), r(H), r(H, H), r(H), r(H, H), r(H), r(H, 4), r(H, H), r(H))
    assert torch.autograd.gradcheck(lambda x, *w: oracle.transformer_conv_ref(x, ei, ea, *w), targs, atol=1e-6)
    nargs = (r(n, H), r(8, 4), r(8), r(H * H, 8), r(H * H), r(H, H), r(H))
    assert torch.autograd.gradcheck(lambda x, *w: oracle.nnconv_mean_ref(x, ei, ea, *w), nargs, atol=1e-6)
    assert torch.autograd.gradcheck(lambda x, *w: oracle.nnconv_mean_factorised_ref(x, ei, ea, *w), nargs, atol=1e-6)
    gargs = (r(8, 5), r(1, 4, 2), r(1, 4, 2), r(8))
    x5 = torch.randn(n, 5, generator=g, dtype=torch.float64)
    assert torch.autograd.gradcheck(lambda *w: oracle.gat_conv_ref(x5, ei, *w), gargs, atol=1e-6)


def test_factorised_nnconv_equals_direct_on_model():
    from gnn_qot_estimation_b200 import synthetic
    hb = synthetic.random_topology_store(200, 700, seed=2).host_batch(0, 1)
    hb.edge_attr = hb.edge_attr.double()
    torch.manual_seed(0)
    a = oracle.TopologicalGNNOracle(200, 32, 3, 4, dropout_p=0.0).double()
    b = oracle.TopologicalGNNOracle(200, 32, 3, 4, dropout_p=0.0, factorised_nnconv=True).double()
    b.load_state_dict(a.state_dict())
    assert torch.allclose(a(hb), b(hb), atol=1e-12)


# ----------------------------------------------------------------------------- (i) checkpoints
@pytest.mark.parametrize("name", ["ckpt_lightpath_model_0.pt", "ckpt_lightpath_model_1.pt"])
def test_lightpath_checkpoints_load_strict(name):
    ck = load_golden(name)
    p = ck["model_params"]
    assert p["in_channels"] == 5 and p["hidden_channels"] == 32 and p["output_dim"] == 3
    assert p["feature_indices"]["is_lut"] == 1
    m = oracle.LightpathGNNOracle(p["in_channels"], p["hidden_channels"], p["output_dim"], is_lut_index=1)
    m.load_state_dict(ck["model_state_dict"], strict=True)
    assert int(m.norm1.module.num_batches_tracked) in (6270, 7315)
    assert sum(v.numel() for v in m.parameters()) == 5507


def test_topological_checkpoint_loads_strict():
    ck = load_golden("ckpt_topological_model_0.pt")
    p = ck["model_params"]
    m = oracle.TopologicalGNNOracle(p["num_nodes"], p["hidden_channels"], p["output_dim"], p["edge_dim"])
    m.load_state_dict(ck["model_state_dict"], strict=True)
    assert sum(v.numel() for v in m.parameters()) == 5291


def test_product_modules_share_state_dict_layout():
    """The drop-in modules must expose exactly the shipped checkpoints' keys/shapes (CPU-side
    construction only; no compute)."""
    from gnn_qot_estimation_b200 import LightpathGNN, TopologicalGNN
    t = TopologicalGNN(75, 16, 3, edge_dim=4)
    sd = load_golden("ckpt_topological_model_0.pt")["model_state_dict"]
    assert {k: tuple(v.shape) for k, v in t.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    t.load_state_dict(sd, strict=True)
    l = LightpathGNN(5, 32, 3, is_lut_index=1)
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    assert {k: (tuple(v.shape), v.dtype) for k, v in l.state_dict().items()} == \
           {k: (tuple(v.shape), v.dtype) for k, v in sd.items()}
    l.load_state_dict(sd, strict=True)


# ----------------------------------------------------------------------------- (iv) golden vectors
def test_golden_lightpath_reproduced():
    g = load_golden("lightpath_eval.pt")
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    b = batch_from_dict(g["batch"])
    m = oracle.LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0).double()
    m.load_state_dict(sd, strict=True)
    m.eval()
    b.x = b.x.double()
    with torch.no_grad():
        out, lb = m(b)
    assert torch.equal(lb, g["expected"]["torch.float64"]["lut_batch"])
    assert rel_err(out, g["expected"]["torch.float64"]["out"]) <= 1e-12
    assert rel_err(g["expected"]["torch.float32"]["out"], out) <= 1e-5


def test_golden_topological_reproduced():
    g = load_golden("topological_train.pt")
    sd = load_golden("ckpt_topological_model_0.pt")["model_state_dict"]
    b = batch_from_dict(g["batch"])
    b.edge_attr = b.edge_attr.double()
    m = oracle.TopologicalGNNOracle(75, 16, 3, 4, dropout_p=0.0).double()
    m.load_state_dict(sd, strict=True)
    out = m(b)
    loss = torch.nn.SmoothL1Loss()(out, b.y.double().view(-1, 3))
    loss.backward()
    e64, e32 = g["expected"]["torch.float64"], g["expected"]["torch.float32"]
    assert rel_err(out, e64["out"]) <= 1e-12 and rel_err(loss, e64["loss"]) <= 1e-12
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert max(grad_errs(grads, e64["grads"], exact_zero=("conv1.lin_key.bias",)).values()) <= 1e-10
    assert max(grad_errs(e32["grads"], grads, exact_zero=("conv1.lin_key.bias",)).values()) <= 1e-5      # the fp32 oracle meets the bar too


# ----------------------------------------------------------------------------- integer side
def test_build_csr_ref_is_stable_argsort():
    ei, _ = _rand_graph(40, 500, 9)
    rp, src, eid = oracle.build_csr_ref(ei, 40)
    order = np.argsort(ei[1].numpy(), kind="stable")
    assert np.array_equal(eid.numpy(), order.astype(np.int32))
    assert np.array_equal(src.numpy(), ei[0].numpy()[order].astype(np.int32))
    assert np.array_equal(np.diff(rp.numpy()), np.bincount(ei[1].numpy(), minlength=40))


def test_gat_edges_ref():
    ei = torch.tensor([[0, 1, 1, 2], [0, 2, 1, 0]])
    out = oracle.gat_edges_ref(ei, 3)
    assert out.tolist() == [[1, 2, 0, 1, 2], [2, 0, 0, 1, 2]]


def test_hand_written_backward_matches_autograd():
    """oracle/topo_fused_math.py (the per-graph forward + SmoothL1 + backward a fused training kernel evaluates,
    written out without autograd) == autograd of TopologicalGNNOracle, fp64, ragged graphs incl. isolated
    nodes and duplicate edges."""
    import torch
    from oracle import TopologicalGNNOracle
    from oracle.topo_fused_math import graph_fwd_bwd
    from gnn_qot_estimation_b200 import Batch
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(6)
    m = TopologicalGNNOracle(20, 16, 3, 4, dropout_p=0.0).double()
    sizes = [14, 3, 20, 1, 9]
    ids, eis, eas, bts, off = [], [], [], [], 0
    per_graph = []
    for gi, n in enumerate(sizes):
        E = 0 if n == 1 else 3 * n
        src, dst = torch.randint(0, n, (E,), generator=g), torch.randint(0, n, (E,), generator=g)
        ea = torch.rand(E, 4, generator=g, dtype=torch.float64)
        nid = torch.randperm(20, generator=g)[:n]
        per_graph.append((nid, src, dst, ea))
        ids.append(nid); eis.append(torch.stack([src, dst]) + off); eas.append(ea); bts.append(torch.full((n,), gi)); off += n
    y = torch.rand(len(sizes), 3, generator=g, dtype=torch.float64) * 3 - 1      # some |diff| > 1: both SmoothL1 branches
    b = Batch(x=None, edge_index=torch.cat(eis, 1), edge_attr=torch.cat(eas), batch=torch.cat(bts),
              node_ids=torch.cat(ids), num_graphs=len(sizes))
    out = m(b)
    loss = torch.nn.SmoothL1Loss()(out, y)
    loss.backward()
    p = {k: v.detach() for k, v in m.state_dict().items()}
    tot, grads = 0.0, {k: torch.zeros_like(v) for k, v in p.items()}
    for gi, (nid, src, dst, ea) in enumerate(per_graph):
        o, l, gg = graph_fwd_bwd(p, nid, src, dst, ea, y[gi], 3 * len(sizes))
        assert torch.allclose(o, out[gi].detach(), rtol=1e-12, atol=1e-12)
        tot = tot + l
        for k, v in gg.items():
            grads[k] += v
    assert abs(float(tot) - float(loss)) < 1e-12
    for k, prm in m.named_parameters():
        assert torch.allclose(grads[k], prm.grad, rtol=1e-9, atol=1e-12), k
