"""Multi-process host logic on CPU (gloo, world_size 2): graph sharding and the flat gradient
all-reduce of gnn_qot_estimation_b200.distributed.  The module under the wrapper is the CPU
oracle (the product modules have no CPU path); what is tested is the wrapper: averaged
per-rank gradients == gradients of the concatenated batch, one flat buffer, parameters
broadcast from rank 0."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_exactly():
    from gnn_qot_estimation_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1_000_000, 10_000_000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    assert len(shard_range(10_000_000, 3, 8)) == 1_250_000          # BASELINE cfg 4
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_flat_grad_buffer_single_process():
    from gnn_qot_estimation_b200.distributed import FlatGradBuffer
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    fb = FlatGradBuffer(m.parameters())
    assert fb.flat.numel() == sum(p.numel() for p in m.parameters())
    m(torch.ones(5, 4)).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()
    fb.gather()
    assert torch.equal(fb.flat, ref)
    for i, p in enumerate(m.parameters()):                  # p.grad now aliases the flat buffer
        assert p.grad.data_ptr() == fb.view_of(i).data_ptr()
    fb.flat.mul_(0.5)
    assert torch.equal(torch.cat([p.grad.reshape(-1) for p in m.parameters()]), ref * 0.5)
    fb.zero()
    assert all(p.grad is None for p in m.parameters())
    m(torch.ones(5, 4)).sum().backward()                    # fresh gradients, not accumulated
    fb.gather()
    assert torch.equal(fb.flat, ref)
    # a parameter that received no gradient contributes zeros
    fb.zero()
    m[0](torch.ones(5, 4)).sum().backward()
    fb.gather()
    n0 = sum(p.numel() for p in m[0].parameters())
    assert torch.equal(fb.flat[:n0], ref.new_tensor([5.0] * 12 + [5.0] * 3)) and float(fb.flat[n0:].abs().sum()) == 0.0


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnn_qot_estimation_b200 import synthetic
        from gnn_qot_estimation_b200.distributed import GraphDataParallel, shard_range
        from oracle import TopologicalGNNOracle
        G = 16
        store = synthetic.nsfnet_store(G, seed=0)
        torch.manual_seed(100 + rank)                         # different init per rank: broadcast must fix it
        model = TopologicalGNNOracle(14, 16, 3, 4, dropout_p=0.0).double()
        ddp = GraphDataParallel(model)
        sd0 = [p.detach().clone() for p in model.parameters()]
        gathered = [None] * world
        dist.all_gather_object(gathered, [p.tolist() for p in sd0])
        assert gathered[0] == gathered[rank], "parameters not broadcast from rank 0"

        def loss_of(m, b):
            b.edge_attr = b.edge_attr.double()
            return torch.nn.SmoothL1Loss()(m(b), b.y.double().view(-1, 3))

        r = shard_range(G, rank, world)
        ddp.zero_grad()
        loss_of(ddp, store.host_batch(r.start, r.stop)).backward()
        ddp.sync_gradients()
        got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        # single-process reference: the whole batch on a copy of the same parameters
        ref_model = TopologicalGNNOracle(14, 16, 3, 4, dropout_p=0.0).double()
        ref_model.load_state_dict(model.state_dict())
        loss_of(ref_model, store.host_batch(0, G)).backward()
        ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        # one optimizer step keeps the replicas identical
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        opt.step()
        after = [None] * world
        dist.all_gather_object(after, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).tolist())
        q.put((rank, err, after[0] == after[rank], len(r)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_equal_concatenated_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, "worker failed (see its traceback above)"
    res = [q.get(timeout=10) for _ in range(world)]
    for rank, err, same, n in res:
        assert n == 8
        assert err <= 1e-12, (rank, err)
        assert same, f"rank {rank} diverged from rank 0 after the optimizer step"
