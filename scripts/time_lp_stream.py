"""Times the persistent multi-batch LightpathGNN eval kernel on resident batches (BASELINE cfg 2 shape):
python scripts/time_lp_stream.py [n_batches] [reps] -> us per 4096-graph batch and the roofline fraction."""
import json, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", map_location="cpu", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
sym = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
store = synthetic.lightpath_store(nb * 4096, seed=1, device=dev)
if sym:
    assert store.verify_layout()
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(nb)]
split = (sys.argv[4] != "0") if len(sys.argv) > 4 else False
plan = m.stream_plan(bs, split_head=split)
for _ in range(3):
    m.forward_stream(plan)
torch.cuda.synchronize()
alg = sum(20 * b.num_nodes + 8 * b.num_edges + 24 * (b.num_graphs + 1) + 24 * r + 4 for b, r in zip(bs, plan.rows)) / nb
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        m.forward_stream(plan)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        g.replay()
    e1.record(s); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * nb)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6533.5
print(json.dumps({"n_batches": nb, "verified_layout": sym, "split_head": split, "us_per_batch": us, "graphs_per_s": 4096 / us * 1e6, "alg_bytes_per_batch": alg,
                  "frac": alg / (us * 1e-6) / 1e9 / peak, "status": int(plan.status.max().item())}))
