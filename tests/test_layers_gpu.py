"""GPU parity of the PyG-named layers in gnn_qot_estimation_b200.nn used ONE BY ONE -- the path a
reference maintainer takes when keeping the reference's models.py and swapping only its layer imports
(topological_training/models.py:3, lightpath_training/models.py:3): each layer is
its own autograd op, the glue (leaky_relu / relu / dropout / boolean LUT mask / torch MLP) stays torch."""
import pytest
import torch
import torch.nn.functional as F

from conftest import grad_errs, load_golden, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_topological_composed_from_layers(cuda):
    from gnn_qot_estimation_b200 import nn as qnn, synthetic
    from oracle import TopologicalGNNOracle
    sd = load_golden("ckpt_topological_model_0.pt")["model_state_dict"]
    o = TopologicalGNNOracle(75, 16, 3, 4, dropout_p=0.0).double()
    o.load_state_dict(sd, strict=True)
    emb = torch.nn.Embedding(75, 16)
    conv1 = qnn.TransformerConv(16, 16, edge_dim=4)
    edge_nn = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 256))
    conv2 = qnn.NNConv(16, 16, nn=edge_nn, aggr="mean")
    mlp = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.LeakyReLU(), torch.nn.Dropout(0.0), torch.nn.Linear(16, 3))
    mods = {"node_embeddings": emb, "conv1": conv1, "conv2": conv2, "mlp": mlp}
    for name, m in mods.items():
        m.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")}, strict=True)
        m.to(cuda)
    hb = synthetic.nsfnet_store(40, seed=2).host_batch(0, 40)
    b = hb.to(cuda)
    x = emb(b.node_ids)
    x = F.leaky_relu(conv1(x, b.edge_index, b.edge_attr))
    x = F.leaky_relu(conv2(x, b.edge_index, b.edge_attr))
    out = mlp(qnn.global_mean_pool(x, b.batch))
    loss = F.smooth_l1_loss(out, b.y.view(-1, 3))
    loss.backward()
    hb.edge_attr = hb.edge_attr.double()
    eo = o(hb)
    el = F.smooth_l1_loss(eo, hb.y.double().view(-1, 3))
    el.backward()
    assert rel_err(out, eo) <= RTOL and rel_err(loss, el) <= RTOL
    got = {f"{n}.{k}": p.grad for n, m in mods.items() for k, p in m.named_parameters()}
    ref = {k: p.grad for k, p in o.named_parameters()}
    for k, e in grad_errs(got, ref, exact_zero=("conv1.lin_key.bias",)).items():
        assert e <= RTOL, (k, e)


@pytest.mark.parametrize("train", [True, False])
def test_lightpath_composed_from_layers(cuda, train):
    from gnn_qot_estimation_b200 import nn as qnn, synthetic
    from oracle import LightpathGNNOracle
    from test_lightpath_train_gpu import _kink_free_batch
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    o = LightpathGNNOracle(5, 32, 3, 1, dropout_p=0.0).double()
    o.load_state_dict(sd, strict=True)
    conv1 = qnn.GATConv(5, 32, heads=4, concat=True)
    norm1 = qnn.BatchNorm(128)
    mlp = torch.nn.Sequential(torch.nn.Linear(128, 32), torch.nn.LeakyReLU(), torch.nn.Dropout(0.0), torch.nn.Linear(32, 3))
    mods = {"conv1": conv1, "norm1": norm1, "mlp": mlp}
    for name, m in mods.items():
        m.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")}, strict=True)
        m.to(cuda).train(train)
    o.train(train)
    hb = _kink_free_batch(o, 60, 1, seed=17)
    b = hb.to(cuda)
    h = F.relu(norm1(conv1(b.x, b.edge_index)))
    mask = b.x[:, 1] == 1.0
    out = mlp(h[mask])
    lut_batch = b.batch[mask]
    loss = F.smooth_l1_loss(out, b.y[lut_batch])
    loss.backward()
    hb.x = hb.x.double()
    eo, elb = o(hb)
    el = F.smooth_l1_loss(eo, hb.y.double()[elb])
    el.backward()
    assert torch.equal(lut_batch.cpu(), elb)
    assert rel_err(out, eo) <= RTOL and rel_err(loss, el) <= RTOL
    got = {f"{n}.{k}": p.grad for n, m in mods.items() for k, p in m.named_parameters()}
    ref = {k: p.grad for k, p in o.named_parameters()}
    zero = ("conv1.bias",) if train else ()
    for k, e in grad_errs(got, ref, exact_zero=zero).items():
        assert e <= RTOL, (k, e)
    if train:
        assert rel_err(norm1.module.running_mean, o.norm1.module.running_mean) <= RTOL
        assert rel_err(norm1.module.running_var, o.norm1.module.running_var) <= RTOL


def test_mean_pool_and_batchnorm_standalone(cuda):
    from gnn_qot_estimation_b200 import nn as qnn
    g = torch.Generator().manual_seed(0)
    sizes = torch.tensor([5, 1, 300, 44])
    x = torch.randn(int(sizes.sum()), 32, generator=g)
    batch = torch.repeat_interleave(torch.arange(4), sizes)
    xd = x.to(cuda).requires_grad_(True)
    p = qnn.global_mean_pool(xd, batch.to(cuda))
    p.square().sum().backward()
    x64 = x.double().requires_grad_(True)
    ref = torch.stack([x64[batch == i].mean(0) for i in range(4)])
    ref.square().sum().backward()
    assert rel_err(p, ref) <= RTOL and rel_err(xd.grad, x64.grad) <= RTOL
    bn = qnn.BatchNorm(32).to(cuda).train()
    bn64 = torch.nn.BatchNorm1d(32).double().train()
    xd2 = x.to(cuda).requires_grad_(True)
    y = bn(xd2)
    (y * torch.arange(32, device=cuda)).sum().backward()
    x642 = x.double().requires_grad_(True)
    y64 = bn64(x642)
    (y64 * torch.arange(32, dtype=torch.float64)).sum().backward()
    assert rel_err(y, y64) <= RTOL
    scale = float(x642.grad.abs().max())
    assert float((xd2.grad.double().cpu() - x642.grad).abs().max()) <= RTOL * max(scale, 1.0)
    assert rel_err(bn.module.running_var, bn64.running_var) <= RTOL


def test_stand_in_batches_feed_the_b200_modules(cuda):
    """torch_geometric stand-in: networkx graphs -> from_networkx -> DataLoader -> Batch.to(cuda) ->
    the drop-in modules; same answers as the oracle on the same batch."""
    import networkx as nx
    from gnn_qot_estimation_b200 import LightpathGNN, TopologicalGNN
    from gnn_qot_estimation_b200.pyg_compat import DataLoader, from_networkx
    from oracle import LightpathGNNOracle, TopologicalGNNOracle
    g = torch.Generator().manual_seed(0)
    items_t, items_l = [], []
    for s in range(6):
        G = nx.gnm_random_graph(12 + s, 30, seed=s)
        d = from_networkx(G)
        d.node_ids = torch.arange(d.num_nodes)
        d.edge_attr = torch.rand(d.edge_index.shape[1], 4, generator=g)
        d.x = None
        d.y = torch.rand(3, generator=g)
        items_t.append(d)
        e = from_networkx(G)
        x = torch.rand(e.num_nodes, 5, generator=g)
        x[:, 1] = 0.0
        x[s % e.num_nodes, 1] = 1.0
        e.x = x
        e.y = torch.rand(1, 3, generator=g)
        items_l.append(e)
    sd = load_golden("ckpt_topological_model_0.pt")["model_state_dict"]
    bt = next(iter(DataLoader(items_t, batch_size=6)))
    m = TopologicalGNN(75, 16, 3, edge_dim=4, dropout_p=0.0); m.load_state_dict(sd); m.to(cuda).eval()
    o = TopologicalGNNOracle(75, 16, 3, 4, dropout_p=0.0).double(); o.load_state_dict(sd); o.eval()
    with torch.no_grad():
        out = m(bt.to(cuda))
        bt.edge_attr = bt.edge_attr.double()
        assert rel_err(out, o(bt)) <= RTOL
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    bl = next(iter(DataLoader(items_l, batch_size=6)))
    m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(cuda).eval()
    o = LightpathGNNOracle(5, 32, 3, 1, dropout_p=0.0).double(); o.load_state_dict(sd); o.eval()
    with torch.no_grad():
        out, lb = m(bl.to(cuda))
        bl.x = bl.x.double()
        eo, el = o(bl)
    assert torch.equal(lb.cpu(), el) and rel_err(out, eo) <= RTOL


def test_fused_smooth_l1_and_metrics(cuda):
    """qot_smooth_l1 == torch SmoothL1Loss (value and gradient), incl. the y[lut_batch] row gather, and
    the device-side accumulator reproduces sklearn's r2_score over several batches."""
    from sklearn.metrics import r2_score
    from gnn_qot_estimation_b200 import ops
    g = torch.Generator().manual_seed(0)
    acc = ops.RegressionMetrics(cuda)
    ys, ps = [], []
    for n in (1, 300, 1025):
        pred = (torch.randn(n, 3, generator=g) * 1.5).to(cuda).requires_grad_(True)
        y = torch.randn(n + 7, 3, generator=g).to(cuda)
        rows = torch.randint(0, n + 7, (n,), generator=g).to(cuda)
        loss = ops.smooth_l1_loss(pred, y, rows, metrics=acc)
        (loss * 2.0).backward()
        p64 = pred.detach().double().cpu().requires_grad_(True)
        ref = torch.nn.SmoothL1Loss()(p64, y.double().cpu()[rows.cpu()])
        (ref * 2.0).backward()
        assert rel_err(loss, ref) <= RTOL and rel_err(pred.grad, p64.grad) <= RTOL
        ys.append(y.cpu()[rows.cpu()]); ps.append(pred.detach().cpu())
    res = acc.result()
    Y, P = torch.cat(ys).numpy(), torch.cat(ps).numpy()
    assert res["count"] == Y.shape[0]
    r2 = r2_score(Y.astype("float64"), P.astype("float64"), multioutput="uniform_average")   # fp64 like the accumulator
    assert abs(res["r2_uniform_average"] - r2) <= 1e-9 * max(1.0, abs(r2))
    # plain form (no row gather) inside a model step
    pred = torch.randn(64, 3, device=cuda, requires_grad=True)
    y = torch.randn(64, 3, device=cuda)
    l1 = ops.smooth_l1_loss(pred, y.view(-1))                 # train.py:112 views y as [-1, 3]
    assert rel_err(l1, torch.nn.SmoothL1Loss()(pred.detach(), y)) <= RTOL
