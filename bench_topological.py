"""Secondary benchmark lines (not the headline): BASELINE configs[2] and configs[4].

  --workload topo_train   TopologicalGNN(14,16,3) fwd + SmoothL1 + bwd + flat NCCL grad all-reduce +
                          SGD(lr .1, momentum .9), 1024 NSFNET graphs per GPU per step (cfg 3)
  --workload topo_stress  TopologicalGNN(10000,256,3) fwd (+bwd) on ONE 10k-node / 80k-directed-edge
                          graph (cfg 5): per-kernel HBM roofline of the fused aggregation kernels
Both print one JSON line on rank 0; `--impl reference` times the CPU oracle on the same step.
"""
from __future__ import annotations

import json
import os
import time

import torch


def _ev():
    return torch.cuda.Event(enable_timing=True)


def run(args):
    if args.workload == "topo_train":
        run_train(args)
    elif args.workload == "lightpath_train":
        run_lightpath_train(args)
    else:
        run_stress(args)


# --------------------------------------------------------------------------- lightpath training step
def run_lightpath_train(args):
    """LightpathGNN train step as lightpath_training/train.py:109-132 drives it (batch 512, SGD(0.1, 0.9),
    SmoothL1 on out vs y[lut_batch]); GAT fwd -> batch-stat BN -> LUT head, and every backward kernel."""
    from gnn_qot_estimation_b200 import LightpathGNN, synthetic
    B = args.batch if args.batch != 4096 else 512
    K, W = min(args.steps, 300), max(3, min(args.warmup, 20))
    crit = torch.nn.SmoothL1Loss()
    if args.impl == "reference":
        from oracle import LightpathGNNOracle
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        m = LightpathGNNOracle(5, 32, 3, 1, dropout_p=0.5).train()
        opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
        hb = synthetic.lightpath_store(B, seed=1).host_batch(0, B)
        steps = min(K, 20)
        for i in range(3 + steps):
            if i == 3:
                t0 = time.perf_counter()
            opt.zero_grad()
            out, lb = m(hb)
            crit(out, hb.y[lb]).backward()
            opt.step()
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "reference", "metric": "lightpath_train_graphs_per_sec", "value": steps * B / dt,
                          "unit": "graphs/s", "steps": steps, "ms_per_step": dt / steps * 1e3,
                          "cpu_baseline": {"kind": "port", "cores": os.cpu_count(), "sample": f"{steps} steps of {B} graphs"}}),
              flush=True)
        return
    print(json.dumps(measure_lightpath_train(torch.device("cuda", 0), B, K, W)), flush=True)


def measure_lightpath_train(dev, B: int = 512, K: int = 300, W: int = 10) -> dict:
    """LightpathGNN train step (lightpath_training/train.py:109-132) at batch B over ragged batches: eager, and through
    the shape-keyed CUDA-graph cache in steady state (every shape already captured)."""
    from gnn_qot_estimation_b200 import LightpathGNN, synthetic
    crit = torch.nn.SmoothL1Loss()
    torch.manual_seed(0)
    model = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.5).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    nb = 16
    store = synthetic.lightpath_store(B * nb, seed=1, device=dev)
    batches = [store.collate(range(i * B, (i + 1) * B)) for i in range(nb)]

    def loss_of(m, b):
        out, lb = m(b)
        return crit(out, b.y[lb])

    def eager(i):
        b = batches[i % nb]
        opt.zero_grad()
        loss = loss_of(model, b)
        loss.backward()
        opt.step()
        return loss.detach()          # a live autograd graph would keep its AccumulateGrad nodes (bound to THIS stream) alive

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = _ev(), _ev()
        e0.record()
        for i in range(n):
            loss = fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, loss
    for i in range(W):
        eager(i)
    eager_ms, _ = timed(eager, min(K, 100))
    # the shape-keyed graph cache: first visit of a batch captures its step, later visits replay it
    from gnn_qot_estimation_b200.graphed import GraphedStepCache
    cache = GraphedStepCache(model, opt, loss_of)
    for i in range(nb):
        cache.step(batches[i])                              # one capture per distinct batch
    ms, loss = timed(lambda i: cache.step(batches[i % nb]), K)
    return {"metric": "lightpath_train_graphs_per_sec", "value": B / (ms * 1e-3), "unit": "graphs/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms, "eager_ms_per_step": eager_ms,
            "captures": cache.captures, "final_loss": float(loss),
            "config": {"workload": f"LightpathGNN train step, batch {B} (ragged), SGD(0.1,0.9), dropout 0.5, "
                                   "one CUDA graph per batch shape (GraphedStepCache), steady state"}}


# --------------------------------------------------------------------------- cfg 3
def run_train(args):
    import torch.distributed as dist
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    from gnn_qot_estimation_b200.distributed import GraphDataParallel
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = 1024
    K, W = min(args.steps, 2000), max(3, min(args.warmup, 50))
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import TopologicalGNNOracle
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        m = TopologicalGNNOracle(14, 16, 3, 4, dropout_p=0.0)
        opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
        hb = synthetic.nsfnet_store(B, seed=0).host_batch(0, B)
        steps = min(K, 30)
        for i in range(3 + steps):
            if i == 3:
                t0 = time.perf_counter()
            opt.zero_grad()
            loss = torch.nn.SmoothL1Loss()(m(hb), hb.y.view(-1, 3))
            loss.backward()
            opt.step()
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "reference", "metric": "topo_train_graphs_per_sec", "value": steps * B / dt,
                          "unit": "graphs/s", "n_gpus": args.gpus, "steps": steps, "ms_per_step": dt / steps * 1e3,
                          "cpu_baseline": {"kind": "port", "cores": os.cpu_count(), "sample": f"{steps} steps of {B} graphs"},
                          "config": {"workload": "BASELINE cfg3 on CPU oracle"}}), flush=True)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = measure_train(world, rank, dev, steps=K, warmup=W, graph=not getattr(args, "no_graph", False))
    if rank == 0:
        print(json.dumps({"metric": "topo_train_graphs_per_sec", "unit": "graphs/s", "n_gpus": world,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic", **res}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_train(world: int, rank: int, dev, steps: int = 300, warmup: int = 20, graph: bool = True,
                  min_timed_ms: float = 60.0, dropout_p: float = 0.0) -> dict:
    """BASELINE cfg 3: TopologicalGNN(14,16,3) train step (zero_grad -> forward -> SmoothL1 -> backward ->
    gradient all-reduce -> SGD(0.1, 0.9)) on 1024 NSFNET graphs per GPU, the process group (if any) already
    initialised.  Returns the JSON fields; every rank must call it (the step holds a collective at N > 1)."""
    import torch.distributed as dist
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    from gnn_qot_estimation_b200.distributed import GraphDataParallel
    B = 1024
    torch.manual_seed(0)
    model = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=dropout_p).to(dev).train()
    ddp = GraphDataParallel(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    # SmoothL1Loss (topological_training/train.py:69) as the fused loss + gradient kernel (qot_smooth_l1)
    from gnn_qot_estimation_b200 import ops
    crit = ops.smooth_l1_loss
    nb = 16
    store = synthetic.nsfnet_store(B * nb, seed=rank).to(dev)
    batches = [store.collate(range(i * B, (i + 1) * B)) for i in range(nb)]

    graphed = None
    if graph:
        # the 16 batches of the shard stay resident in HBM and come back every 16 steps (as the chunks of
        # train.py:77-95 do): one captured graph per batch, replayed on the batch's own tensors -- no input copies
        from gnn_qot_estimation_b200.graphed import GraphedStepCache
        graphed = GraphedStepCache(model, opt, lambda m, b: crit(m(b), b.y.view(-1, 3)), ddp=ddp, borrow_inputs=True)
        for b in batches:
            graphed.step(b)

    def step(i):
        b = batches[i % nb]
        if graphed is not None:
            return graphed.step(b)                       # whole step = one CUDA-graph replay
        ddp.zero_grad()
        loss = crit(ddp(b), b.y.view(-1, 3))
        loss.backward()
        ddp.sync_gradients()
        opt.step()
        return loss

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(n):
        e0, e1 = _ev(), _ev()
        sync()
        prof = bool(os.environ.get("QOT_PROFILE_TIMED_REGION"))   # `ncu --profile-from-start off`: only the timed steps
        if prof:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(n):
            loss = step(i)
        e1.record()
        sync()
        if prof:
            torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), loss

    for i in range(max(warmup, 3)):
        step(i)
    ms, loss = timed(steps)
    n = steps
    if ms < min_timed_ms:                                    # every rank sees the same (max-reduced) time
        n = int(steps * min(50.0, min_timed_ms / max(ms, 1e-3))) + 1
        ms, loss = timed(n)
    out = {"value": world * n * B / (ms * 1e-3), "steps": n, "warmup": max(warmup, 3), "ms_per_step": ms / n,
           "final_loss": float(loss),
           "config": {"workload": "BASELINE cfg3: TopologicalGNN(14,16,3) train step, NSFNET graphs, "
                                  f"batch {B}/GPU, SGD(0.1, 0.9), dropout {dropout_p}, flat grad all-reduce x{world}",
                      "cuda_graph": graphed is not None,
                      "exchange": (next(iter(graphed.entries.values())).exchange if graphed is not None
                                   else ("eager" if world > 1 else "none")),
                      "inputs": "resident batches replayed in place (one captured graph per batch, no input copies)"}}
    # the collective alone: the flat gradient all-reduce as the step issues it, back to back
    if world > 1:
        e0, e1 = _ev(), _ev()
        for _ in range(10):
            ddp.grads.all_reduce_mean()
        sync()
        e0.record()
        for _ in range(100):
            ddp.grads.all_reduce_mean()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["allreduce_us"] = float(t) * 10.0                # ms / 100 calls -> us
        out["allreduce_bytes"] = int(ddp.grads.flat.numel() * 4)
    return out


# --------------------------------------------------------------------------- cfg 5
def run_stress(args):
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    N, LINKS, H = 10000, 40000, 256
    if args.impl == "reference":
        from oracle import TopologicalGNNOracle
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        m = TopologicalGNNOracle(N, H, 3, 4, dropout_p=0.0, factorised_nnconv=True).eval()
        hb = synthetic.random_topology_store(N, LINKS, seed=2).host_batch(0, 1)
        with torch.no_grad():
            m(hb)
            t0 = time.perf_counter()
            for _ in range(5):
                m(hb)
            dt = (time.perf_counter() - t0) / 5
        print(json.dumps({"impl": "reference", "metric": "topo_stress_fwd_ms", "value": dt * 1e3, "unit": "ms",
                          "cpu_baseline": {"kind": "port", "cores": os.cpu_count(),
                                           "sample": "5 forwards, factorised NNConv (direct form needs 21 GB)"}}), flush=True)
        return
    dev = torch.device("cuda", 0)
    res = measure_stress(dev)
    print(json.dumps({"metric": "topo_stress_fwd_ms", "unit": "ms", "n_gpus": 1, "higher_is_better": False,
                      "dtype": "f32", "data": "synthetic", **res}), flush=True)


def measure_stress(dev, reps: int = 20) -> dict:
    """BASELINE cfg 5: TopologicalGNN(10000,256,3) forward (and forward + backward) on ONE 10k-node /
    80k-directed-edge graph, L2 flushed between iterations; per-kernel roofline of the two aggregation kernels
    (algorithmic bytes of SURVEY 8d / BASELINE.md section 4 over their CUDA-event time)."""
    from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
    N, LINKS, H = 10000, 40000, 256
    torch.manual_seed(0)
    model = TopologicalGNN(N, H, 3, edge_dim=4, dropout_p=0.0).to(dev)
    b = synthetic.random_topology_store(N, LINKS, seed=2).to(dev).collate(range(0, 1))
    E = b.num_edges
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def timed(fn, reps=reps):
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = _ev(), _ev()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model(b)
        fwd_ms = timed(lambda: model(b))
        # the same forward as ONE CUDA-graph replay (~25 launches of 3-70 us each: a third of the eager time is launch gaps)
        fwd_graph_ms = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model(b)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    out_g = model(b)
                g.replay()
                torch.cuda.synchronize()
            torch.cuda.current_stream().wait_stream(side)
            assert torch.equal(out_g, model(b))
            fwd_graph_ms = timed(lambda: g.replay())
        except Exception as e:                                  # noqa: BLE001 -- the eager figure stands
            fwd_graph_ms = f"capture failed: {type(e).__name__}"
    model.train()

    def fb():
        model.zero_grad(set_to_none=True)
        torch.nn.SmoothL1Loss()(model(b), b.y.view(-1, 3)).backward()
    for _ in range(3):
        fb()
    fb_ms = timed(fb)
    conv_bytes = 8 * H * N + 4 * (N + 1) + 4 * E + 16 * E
    return {"value": fwd_ms, "fwd_cuda_graph_ms": fwd_graph_ms, "fwd_bwd_ms": fb_ms,
            "alg_bytes": {"conv_fwd_each": conv_bytes, "fwd_total": 2 * conv_bytes + 4 * H * N + 4 * 2 + 12},
            "config": {"workload": f"BASELINE cfg5: TopologicalGNN({N},{H},3), one graph, {N} nodes, {E} directed "
                                   "edges; L2 flushed between iterations"}}
