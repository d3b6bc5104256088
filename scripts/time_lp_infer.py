"""Developer tool: resident-batch step time of qot_lightpath_infer for the active variant
(QOT_LP_VARIANT) with S graph branches, without bench.py's e2e / CPU legs.
usage: python scripts/time_lp_infer.py [streams] [nbatch] [batch]"""
import sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nbatch = int(sys.argv[2]) if len(sys.argv) > 2 else 96
BS = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
store = synthetic.lightpath_store(nbatch * BS, seed=1, device=dev)
bs = [store.collate(range(i * BS, (i + 1) * BS)) for i in range(nbatch)]
outs = [m.forward_device(b) for b in bs]
torch.cuda.synchronize()
side = torch.cuda.Stream()
streams = [torch.cuda.Stream() for _ in range(S)]


def run():
    cur = torch.cuda.current_stream()
    for b in streams:
        b.wait_stream(cur)
    for i in range(nbatch):
        with torch.cuda.stream(streams[i % S]):
            m.forward_device(bs[i], outs[i])
    for b in streams:
        cur.wait_stream(b)


with torch.cuda.stream(side):
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        run()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    ev0.record(side)
    for _ in range(50):
        g.replay()
    ev1.record(side)
    torch.cuda.synchronize()
print(f"variant {m.dominant_kernel} streams={S} batch={BS}: {ev0.elapsed_time(ev1) * 1e3 / (50 * nbatch):.3f} us per step")
