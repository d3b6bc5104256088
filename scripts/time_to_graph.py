"""Developer tool: device graph construction throughput (both representations), kernels timed with CUDA
events around the whole call (scan launch + offset scan + one host sync + pack launch)."""
import sys, time, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import synthetic
from gnn_qot_estimation_b200.to_graph import create_lightpath_graphs, create_topological_graphs
dev = torch.device("cuda:0")
for (L, Q) in ((60, 80), (37, 50)):
    s = synthetic.network_status_samples(64, L, Q, seed=1)
    data = torch.from_numpy(s["data"]).to(dev).repeat(64, 1, 1, 1)      # 4096 samples
    tgt = torch.from_numpy(s["target"]).repeat(64, 1).to(dev); fr = torch.from_numpy(s["freqs"]).to(dev)
    for name, fn in (("lightpath", lambda: create_lightpath_graphs(data, tgt, fr, s["lp_feat"], s["metric"])),
                     ("topological", lambda: create_topological_graphs(data, tgt, s["lp_feat"], s["metric"]))):
        for _ in range(3):
            st = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            st = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gb = data.numel() * 4 / 1e9                                       # the sample tensor is read once
        print(f"{name:12s} [F=10,L={L},Q={Q}] x {data.shape[0]} samples: {ms:.3f} ms = {data.shape[0]/ms*1e3:.3g} samples/s, "
              f"{gb/ms*1e3:.0f} GB/s of input (read once), nodes {int(st.node_ptr[-1])}, edges {int(st.edge_ptr[-1])}")
