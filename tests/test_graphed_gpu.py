"""GPU: the CUDA-graphed training step (gnn_qot_estimation_b200.graphed) takes exactly the same
optimisation trajectory as the eager step, and the pooling kernels split large graphs correctly."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_graphed_step_equals_eager(cuda):
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    from gnn_qot_estimation_b200.graphed import GraphedTrainStep
    torch.manual_seed(0)
    m1 = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(cuda)
    m2 = copy.deepcopy(m1)
    crit = torch.nn.SmoothL1Loss()
    o1 = torch.optim.SGD(m1.parameters(), lr=0.1, momentum=0.9)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9)
    store = synthetic.nsfnet_store(64 * 6, seed=3).to(cuda)
    batches = [store.collate(range(i * 64, (i + 1) * 64)) for i in range(6)]
    g = GraphedTrainStep(m2, o2, crit, batches[0], warmup=2)
    # constructing the step must not move the model or the optimizer state (snapshot / restore inside)
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p1, p2), k
    for b in batches:
        o1.zero_grad()
        l1 = crit(m1(b), b.y.view(-1, 3))
        l1.backward()
        o1.step()
        l2 = g.step(b)
        assert torch.equal(l1.detach(), l2), (float(l1), float(l2))
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p1, p2), k
    with pytest.raises(RuntimeError, match="static shapes"):
        g.step(store.collate(range(0, 32)))


def test_graphed_step_follows_lr_schedule(cuda):
    """StepLR changes param_groups['lr'] between steps (topological_training/train.py:67,152): the captured
    update must follow it, and a resumed optimizer (existing momentum) must keep its state."""
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    from gnn_qot_estimation_b200.graphed import GraphedTrainStep
    torch.manual_seed(1)
    m1 = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(cuda)
    crit = torch.nn.SmoothL1Loss()
    o1 = torch.optim.SGD(m1.parameters(), lr=0.1, momentum=0.9)
    store = synthetic.nsfnet_store(32 * 6, seed=4).to(cuda)
    batches = [store.collate(range(i * 32, (i + 1) * 32)) for i in range(6)]
    # one eager step first: the graphed optimizer starts from existing momentum buffers
    o1.zero_grad(); crit(m1(batches[0]), batches[0].y.view(-1, 3)).backward(); o1.step()
    m2 = copy.deepcopy(m1)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9)
    o2.load_state_dict(copy.deepcopy(o1.state_dict()))
    s1 = torch.optim.lr_scheduler.StepLR(o1, step_size=2, gamma=0.5)
    s2 = torch.optim.lr_scheduler.StepLR(o2, step_size=2, gamma=0.5)
    g = GraphedTrainStep(m2, o2, crit, batches[0], warmup=2)
    for b in batches:
        o1.zero_grad()
        l1 = crit(m1(b), b.y.view(-1, 3))
        l1.backward()
        o1.step(); s1.step()
        l2 = g.step(b); s2.step()
        assert torch.equal(l1.detach(), l2), (float(l1), float(l2))
    assert o1.param_groups[0]["lr"] == o2.param_groups[0]["lr"] == 0.1 * 0.5 ** 3
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p1, p2), k


def test_ragged_lightpath_training_through_the_graph_cache(cuda):
    """LightpathGNN training (lightpath_training/train.py:109-132: out, lut_batch = model(data); loss on y[lut_batch];
    SGD) over RAGGED batches through GraphedStepCache: one captured graph per batch shape, replayed when the shape
    comes back (train.py:77-95 revisits its chunks with shuffle=False).  Same trajectory as the eager loop, bit for
    bit, including BatchNorm running statistics; no device->host read inside a step."""
    from gnn_qot_estimation_b200 import LightpathGNN, synthetic
    from gnn_qot_estimation_b200.graphed import GraphedStepCache
    torch.manual_seed(2)
    m1 = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0).to(cuda).train()
    m2 = copy.deepcopy(m1)
    crit = torch.nn.SmoothL1Loss()
    o1 = torch.optim.SGD(m1.parameters(), lr=0.1, momentum=0.9)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9)
    store = synthetic.lightpath_store(5 * 96, seed=6).to(cuda)
    batches = [store.collate(range(i * 96, (i + 1) * 96)) for i in range(5)]
    assert all(b.lut_rows == 96 for b in batches) and len({b.num_nodes for b in batches}) > 1   # ragged

    def loss_of(model, b):
        out, lb = model(b)
        return crit(out, b.y[lb])
    cache = GraphedStepCache(m2, o2, loss_of, borrow_inputs=True)    # replay on the batches' own tensors: no input copies
    for epoch in range(3):                                   # epoch 0 captures, epochs 1-2 replay
        for b in batches:
            o1.zero_grad()
            l1 = loss_of(m1, b)
            l1.backward()
            o1.step()
            l2 = cache.step(b)
            assert torch.equal(l1.detach(), l2), (epoch, float(l1), float(l2))
    assert cache.captures == 5 and cache.replays == 15
    for (k, p1), (_, p2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(p1, p2), k


@pytest.mark.parametrize("sizes", [[3000], [700, 1, 1300, 40], [260] * 7])
def test_pool_splits_large_graphs(cuda, sizes):
    """global_mean_pool + head on graphs large enough to be split across blocks (stage-1 partials)."""
    from gnn_qot_estimation_b200 import ops
    from oracle import global_mean_pool_ref
    H = 64
    g = torch.Generator().manual_seed(1)
    N = sum(sizes)
    x = torch.randn(N, H, generator=g)
    gptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    W1, b1 = torch.randn(H, H, generator=g) * 0.1, torch.randn(H, generator=g)
    W2, b2 = torch.randn(3, H, generator=g) * 0.1, torch.randn(3, generator=g)
    xd = x.to(cuda).requires_grad_(True)
    ps = [t.to(cuda).requires_grad_(True) for t in (W1, b1, W2, b2)]
    out = ops.pool_mlp(xd, gptr.to(cuda), *ps)
    out.square().sum().backward()
    x64 = x.double().requires_grad_(True)
    q = [t.double().requires_grad_(True) for t in (W1, b1, W2, b2)]
    pooled = global_mean_pool_ref(x64, batch, len(sizes))
    ref = torch.nn.functional.linear(torch.nn.functional.leaky_relu(torch.nn.functional.linear(pooled, q[0], q[1])), q[2], q[3])
    ref.square().sum().backward()
    assert rel_err(out, ref) <= 1e-5
    assert rel_err(xd.grad, x64.grad) <= 1e-5
    for a, b in zip(ps, q):
        assert rel_err(a.grad, b.grad) <= 1e-5
