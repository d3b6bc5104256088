#!/usr/bin/env python
"""Benchmark of the B200 message-passing hot path (contract: see DESIGN.md "Measurement").

Default workload = BASELINE.json configs[1]: LightpathGNN inference (shipped weights
lightpath_training/models/model_1.pth) over 1,000,000 synthetic lightpath graphs per GPU in
batches of 4096.  One STEP = one pass of the fused eval kernel chain over one batch.

  value   graphs/s, whole job, batches already resident in HBM (reference tensor layout:
          fp32 x, int64 edge_index), steps replayed from a CUDA graph, timed with CUDA events
  e2e     graphs/s through the public host-facing API (LightpathInferencePipeline): pinned HOST
          batches -> H2D -> kernels -> D2H of (out, lut_batch) every step
  roofline  algorithmic bytes of the dominant kernel / its measured duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (pure-PyTorch port of the reference's PyG path) on the host cores

`--impl reference` times that CPU oracle alone, all host threads, same config and metric.
N>1 (torchrun): every rank owns a disjoint shard of graphs (weak scaling, no collective on the
data path); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

CKPT = ROOT / "tests" / "golden" / "ckpt_lightpath_model_1.pt"
METRIC = "lightpath_infer_graphs_per_sec"
UNIT = "graphs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24500)
    ap.add_argument("--warmup", type=int, default=245)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graphs", type=int, default=1_000_000, help="graphs per GPU shard")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=490)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=4, help="batches in flight in the host pipeline")
    ap.add_argument("--no-graph", action="store_true", help="topo_train: eager step instead of the CUDA-graphed step")
    ap.add_argument("--workload", default="lightpath_infer", choices=["lightpath_infer", "topo_train", "topo_stress", "lightpath_train"],
                    help="lightpath_infer = BASELINE configs[1] (the headline line); topo_train = configs[2] "
                         "(TopologicalGNN DDP training, batch 1024/GPU); topo_stress = configs[4] (10k nodes, hidden 256)")
    ap.add_argument("--streams", type=int, default=16,
                    help="independent batches in flight in the resident run (graph branches)")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while a timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int, period: float = 0.01):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------- helpers
def load_state():
    return torch.load(CKPT, map_location="cpu", weights_only=False)["model_state_dict"]


def algorithmic_bytes(batch) -> int:
    """Compulsory traffic of the fused eval kernel for one batch ('lp_infer'):
    x (20 B/node) + destination row of edge_index (8 B/edge) + gptr/eptr/lut_ptr (24 B/graph) +
    out/lut_batch/lut_node rows (24 B/LUT row, L = lut_ptr[B])."""
    N, E, B = batch.num_nodes, batch.num_edges, batch.num_graphs
    L = B if batch.lut_ptr is None else int(batch.lut_ptr[-1])
    return 20 * N + 8 * E + 24 * (B + 1) + 24 * L + 4


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_oracle_rate(host_batches, seconds: float, threads: int):
    """graphs/s of the CPU oracle (oracle.LightpathGNNOracle, fp32, eval) on a bounded sample."""
    from oracle import LightpathGNNOracle
    torch.set_num_threads(threads)
    m = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    m.load_state_dict(load_state(), strict=True)
    m.eval()
    done, t_used, i = 0, 0.0, 0
    with torch.no_grad():
        m(host_batches[0])                                   # warm-up
        while t_used < seconds:
            b = host_batches[i % len(host_batches)]
            t0 = time.perf_counter()
            m(b)
            t_used += time.perf_counter() - t0
            done += b.num_graphs
            i += 1
    return done / t_used, i, done


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gnn_qot_estimation_b200 import synthetic
    threads = os.cpu_count() or 1
    n_b = 4
    store = synthetic.lightpath_store(args.batch * n_b, seed=1, device="cpu")
    hbs = [store.host_batch(i * args.batch, (i + 1) * args.batch) for i in range(n_b)]
    from oracle import LightpathGNNOracle
    torch.set_num_threads(threads)
    m = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    m.load_state_dict(load_state(), strict=True)
    m.eval()
    steps = max(1, min(args.steps, 200))                      # bounded: each step = one 4096-graph batch
    warm = max(1, min(args.warmup, 3))
    with torch.no_grad():
        for i in range(warm):
            m(hbs[i % n_b])
        t0 = time.perf_counter()
        g = 0
        for i in range(steps):
            m(hbs[i % n_b])
            g += hbs[i % n_b].num_graphs
        dt = time.perf_counter() - t0
    v = g / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE cfg2: LightpathGNN eval, synthetic lightpath graphs n~U{8..56}, batch 4096",
                   "batch": args.batch, "weights": "lightpath_training/models/model_1.pth"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} batches of {args.batch} graphs, pure-PyTorch oracle (PyG absent)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    from gnn_qot_estimation_b200 import LightpathGNN, synthetic
    from gnn_qot_estimation_b200.pipeline import LightpathInferencePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from gnn_qot_estimation_b200.distributed import bind_to_gpu_numa_node
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = bind_to_gpu_numa_node(local)          # before any pinned allocation: host buffers land next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0)
    model.load_state_dict(load_state(), strict=True)
    model.to(dev).eval()

    # ---- shard of graphs owned by this rank, generated on the device (seed 1 + rank)
    G, Bsz = args.graphs, args.batch
    store = synthetic.lightpath_store(G, seed=1 + rank, device=dev)
    nb = (G + Bsz - 1) // Bsz
    batches = [store.collate(range(i * Bsz, min((i + 1) * Bsz, G))) for i in range(nb)]
    outs = [model.forward_device(b) for b in batches]            # eager pass: allocates outputs, warms up
    torch.cuda.synchronize()
    input_bytes = sum(b.nbytes(("x", "edge_index", "ptr", "edge_ptr", "lut_ptr")) for b in batches)
    launches_per_step = model.launches_per_step

    K, W = args.steps, max(args.warmup, 3)
    S = max(1, args.streams)
    side = torch.cuda.Stream()
    branches = [torch.cuda.Stream() for _ in range(S)] if S > 1 else [side]

    def run_steps(first, count, fork=False):
        """`count` steps starting at batch `first`; with S > 1 consecutive steps go round-robin onto S
        streams forked from / joined back into the current one (independent batches overlap)."""
        cur = torch.cuda.current_stream()
        if fork and S > 1:
            for b in branches:
                b.wait_stream(cur)
        for s in range(first, first + count):
            i = s % nb
            if fork and S > 1:
                with torch.cuda.stream(branches[s % S]):
                    model.forward_device(batches[i], outs[i])
            else:
                model.forward_device(batches[i], outs[i])
        if fork and S > 1:
            for b in branches:
                cur.wait_stream(b)

    # warm-up (eager, on the streams that will be captured so their scratch exists), then capture
    # the K timed steps as CUDA graphs: whole passes over the shard + a remainder
    passes, rem = divmod(K, nb)
    g_pass = g_rem = None
    with torch.cuda.stream(side):
        run_steps(0, W, fork=True)
        torch.cuda.synchronize()
        if passes:
            g_pass = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_pass, stream=side):
                run_steps(0, nb, fork=True)
        if rem:
            g_rem = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_rem, stream=side):
                run_steps(0, rem, fork=True)
        if g_pass is not None:
            g_pass.replay()
        torch.cuda.synchronize()
        graphs_done = passes * G + sum(batches[i].num_graphs for i in range(rem))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with ClockSampler(local) as clk:
            ev0.record(side)
            for _ in range(passes):
                g_pass.replay()
            if g_rem is not None:
                g_rem.replay()
            ev1.record(side)
            torch.cuda.synchronize()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * graphs_done / (ms_max * 1e-3)

    # ---- checksum of the last pass against an eager run through the oracle-checked module path
    n_chk = min(nb, 3)
    for i in range(n_chk):
        with torch.no_grad():
            ref = model(batches[i])[0]
        n = int(outs[i].n_lut.item())
        assert n == ref.shape[0] and torch.equal(outs[i].out[:n], ref), "graph replay diverged from eager path"

    # ---- roofline of the dominant kernel: duration measured live (steps are back to back on one
    # stream; the step is the kernel chain of qot_lightpath_infer)
    alg = sum(algorithmic_bytes(b) for b in batches) / nb
    step_s = ms * 1e-3 / K
    peak, peak_kind = peaks()
    achieved = alg / step_s / 1e9
    traffic = None
    tp = ROOT / "profiles" / "r1_lp_infer_traffic.json"          # dram__bytes_{read,write}.sum of one ncu --set full capture
    if tp.exists() and Bsz == 4096:
        tj = json.loads(tp.read_text())
        traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": model.dominant_kernel, "alg_bytes_per_launch": alg,
                "avg_launch_us": step_s * 1e6, "launches_in_flight": S, "peak_kind": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)"}

    # ---- end to end: pinned host batches -> H2D -> kernels -> D2H, through the public pipeline API
    e2e = None
    if not args.no_e2e:
        n_host = min(nb, 64)
        cpu_store = synthetic.lightpath_store(n_host * Bsz, seed=101 + rank, device=dev)
        cpu_store = cpu_store.to("cpu")
        hbs = [cpu_store.host_batch(i * Bsz, (i + 1) * Bsz, pin=True) for i in range(n_host)]
        pipe = LightpathInferencePipeline(model, max_nodes=max(b.num_nodes for b in hbs),
                                          max_edges=max(b.num_edges for b in hbs), max_graphs=Bsz, depth=args.e2e_depth)
        Ke = max(1, min(args.e2e_steps, K))
        seq = [hbs[i % n_host] for i in range(Ke)]
        pipe.run(seq[: max(3, min(W, 16))])                    # warm-up
        barrier()
        t0 = time.perf_counter()
        res = pipe.run(seq)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ge = sum(b.num_graphs for b in seq)
        # parity of the pipeline against the module path on one batch
        with torch.no_grad():
            o_ref, l_ref = model(hbs[0].to(dev))
        assert torch.equal(res[0][0], o_ref.cpu()) and torch.equal(res[0][1], l_ref.cpu())
        e2e = {"value": world * ge / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": (pipe.h2d_bytes + pipe.zero_copy_bytes) / max(pipe.steps, 1),
               "d2h_bytes_per_step": pipe.d2h_bytes / max(pipe.steps, 1), "steps": Ke,
               "h2d_note": "copied in ONE transfer per step (the host batch keeps them contiguous): destination row of "
                           "edge_index, ptr/edge_ptr/lut_ptr, x; the source row stays in "
                           "pinned host memory and the kernel reads ~4 sectors (32 B) per LUT row from it over PCIe "
                           f"(~{pipe.zero_copy_bytes / max(pipe.steps, 1):.0f} B/step, estimated, included)"}

    cpu_base = None
    if all_cpus is not None:
        os.sched_setaffinity(0, all_cpus)        # the CPU baseline gets every host core back
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cs = synthetic.lightpath_store(4 * Bsz, seed=1, device="cpu")
        chb = [cs.host_batch(i * Bsz, (i + 1) * Bsz) for i in range(4)]
        v, n_it, n_g = cpu_oracle_rate(chb, args.cpu_seconds, threads)
        cpu_base = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{n_it} batches of {Bsz} graphs ({n_g} graphs), pure-PyTorch oracle of the "
                              f"PyG path (torch_geometric absent), fp32, {threads} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE cfg2: LightpathGNN eval (GAT->BN->ReLU->LUT->MLP), "
                                   f"{G} synthetic lightpath graphs per GPU, n~U{{8..56}}, batch {Bsz}",
                       "batch": Bsz, "graphs_per_gpu": G, "weights": "lightpath_training/models/model_1.pth",
                       "l2": f"inputs cycle through {input_bytes / 1e9:.2f} GB of distinct batches (> 126 MB L2)",
                       "parallelism": f"graph-sharded x{world}, no collective", "host": numa},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e,
            "gpu_launches": launches_per_step * K, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.workload != "lightpath_infer":
        import bench_topological
        bench_topological.run(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
