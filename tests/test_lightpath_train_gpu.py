"""GPU parity: LightpathGNN training path (GATConv -> BatchNorm with batch statistics ->
ReLU -> LUT readout -> MLP; lightpath_training/models.py:26-45 under model.train()) and its
backward, as driven by lightpath_training/train.py:114-132 (SmoothL1 on out vs y[lut_batch]).
Tolerance: 1e-5 relative (BASELINE.json north_star), written below as RTOL."""
import pytest
import torch

from conftest import grad_parity, load_golden, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _pair(dev, sd=None, seed=0):
    from gnn_qot_estimation_b200 import LightpathGNN
    from oracle import LightpathGNNOracle
    torch.manual_seed(seed)
    m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    else:
        with torch.no_grad():
            m.conv1.bias.normal_(0, 0.1)
            m.norm1.module.weight.uniform_(0.5, 1.5)
            m.norm1.module.bias.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    o = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0).double()
    o.load_state_dict(sd, strict=True)
    o32 = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    o32.load_state_dict(sd, strict=True)
    return m.to(dev), o, o32


def _step(model, b, dtype=None):
    model.zero_grad(set_to_none=True)
    out, lut_batch = model(b)
    y = b.y[lut_batch]                              # lightpath_training/train.py:122-123
    if dtype is not None:
        y = y.to(dtype)
    loss = torch.nn.SmoothL1Loss()(out, y)
    loss.backward()
    return out.detach(), lut_batch, loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def _b64(b):
    bb = b.to("cpu")
    bb.x = bb.x.double()
    return bb


def _kink_margin(o64, hb):
    """Smallest |pre-activation| feeding a ReLU / LeakyReLU that the loss depends on (LUT rows).
    The derivative of those jumps at 0, so a value within fp32 round-off of 0 makes the GRADIENT of
    any fp32 implementation differ from fp64 by a discrete amount -- such inputs cannot carry a
    1e-5 gradient comparison (observed: the fp32 oracle itself lands on either side depending on
    the host CPU)."""
    import copy
    o = copy.deepcopy(o64)
    b = _b64(hb)
    with torch.no_grad():
        h = o.norm1(o.conv1(b.x, b.edge_index))
        pre1 = h[b.x[:, 1] == 1.0]
        pre2 = o.mlp[0](torch.relu(pre1))
    return min(float(pre1.abs().min()), float(pre2.abs().min()))


def _kink_free_batch(o64, num_graphs, luts, seed):
    from gnn_qot_estimation_b200 import synthetic
    for attempt in range(20):
        hb = synthetic.lightpath_store(num_graphs, seed=seed + 1000 * attempt, lut_per_graph=luts).host_batch(0, num_graphs)
        if _kink_margin(o64, hb) > 2e-6:          # ~10x the fp32 round-off of O(1) pre-activations
            return hb
    raise AssertionError("no kink-free batch found")


@pytest.mark.parametrize("ckpt", ["ckpt_lightpath_model_1.pt", None])
@pytest.mark.parametrize("num_graphs,luts", [(4, 1), (96, 1), (300, 2)])
def test_train_step_vs_oracle(cuda, ckpt, num_graphs, luts):
    from gnn_qot_estimation_b200 import synthetic
    sd = load_golden(ckpt)["model_state_dict"] if ckpt else None
    m, o, o32 = _pair(cuda, sd, seed=num_graphs)
    m.train(); o.train(); o32.train()
    hb = _kink_free_batch(o, num_graphs, luts, seed=2 + num_graphs)
    out, lb, loss, grads = _step(m, hb.to(cuda))
    eo, el, eloss, eg = _step(o, _b64(hb), torch.float64)
    _, _, _, eg32 = _step(o32, hb.to("cpu"))
    assert torch.equal(lb.cpu(), el)
    assert rel_err(out, eo) <= RTOL and rel_err(loss, eloss) <= RTOL
    # conv1.bias feeds a batch-statistics BatchNorm: its exact gradient is 0
    grad_parity(grads, eg, eg32, RTOL, exact_zero=("conv1.bias",))
    # running statistics (momentum 0.1, unbiased variance) and the batch counter
    bn, obn = m.norm1.module, o.norm1.module
    assert rel_err(bn.running_mean, obn.running_mean) <= RTOL
    assert rel_err(bn.running_var, obn.running_var) <= RTOL
    assert int(bn.num_batches_tracked) == int(obn.num_batches_tracked)


def test_eval_with_grad_uses_running_stats(cuda):
    """eval() but grad enabled (fine-tuning a frozen-BN model): general path, running stats."""
    from gnn_qot_estimation_b200 import synthetic
    sd = load_golden("ckpt_lightpath_model_0.pt")["model_state_dict"]
    m, o, o32 = _pair(cuda, sd)
    m.eval(); o.eval(); o32.eval()
    hb = _kink_free_batch(o, 50, 1, seed=3)
    out, lb, loss, grads = _step(m, hb.to(cuda))
    eo, el, eloss, eg = _step(o, _b64(hb), torch.float64)
    _, _, _, eg32 = _step(o32, hb.to("cpu"))
    assert torch.equal(lb.cpu(), el) and rel_err(out, eo) <= RTOL
    grad_parity(grads, eg, eg32, RTOL)


def test_general_path_equals_fused_eval_path(cuda):
    from gnn_qot_estimation_b200 import synthetic
    sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
    m, _, _ = _pair(cuda, sd)
    m.eval()
    b = synthetic.lightpath_store(200, seed=6).host_batch(0, 200).to(cuda)
    with torch.no_grad():
        o1, l1 = m._forward_eval(b)
        o2, l2 = m._forward_general(b)
    assert torch.equal(l1, l2)
    assert rel_err(o1, o2) <= RTOL


def test_train_no_lut_raises(cuda):
    from gnn_qot_estimation_b200 import synthetic
    m, _, _ = _pair(cuda)
    m.train()
    hb = synthetic.lightpath_store(4, seed=1).host_batch(0, 4)
    hb.x[:, 1] = 0.5
    hb.lut_ptr = None
    with pytest.raises(ValueError, match="No LUT node found in the batch."):
        m(hb.to(cuda))


def test_train_deterministic(cuda):
    from gnn_qot_estimation_b200 import synthetic
    m, _, _ = _pair(cuda, seed=3)
    m.train()
    b = synthetic.lightpath_store(512, seed=9).host_batch(0, 512).to(cuda)
    o1, _, l1, g1 = _step(m, b)
    o2, _, l2, g2 = _step(m, b)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k
