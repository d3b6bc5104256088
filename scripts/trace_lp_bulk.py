"""Developer tool: per-block phase durations (SM cycles) of lp_infer_bulk_kernel under the bench's
concurrency (8 streams in flight).  Needs a -DQOT_LP_TRACE build: scripts/build_trace.sh, then
QOT_B200_LIB=scripts/libqot_b200_trace.so python scripts/trace_lp_bulk.py [streams]"""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic, _lib
dev = torch.device("cuda:0")
nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
nbatch = 64
store = synthetic.lightpath_store(nbatch * 4096, seed=1, device=dev)
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(nbatch)]
outs = [m.forward_device(b) for b in bs]
torch.cuda.synchronize()
L = _lib.lib()
L.qot_debug_set_lp_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(600 * 8, dtype=torch.int64, device=dev)
assert L.qot_debug_set_lp_trace(trace.data_ptr()) == 0
side = torch.cuda.Stream()
streams = [torch.cuda.Stream() for _ in range(nstreams)]


def run():
    cur = torch.cuda.current_stream()
    for b in streams:
        b.wait_stream(cur)
    for i in range(nbatch):
        with torch.cuda.stream(streams[i % nstreams]):
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
    for _ in range(20):
        g.replay()
    ev1.record(side)
    torch.cuda.synchronize()
print(f"{ev0.elapsed_time(ev1) * 1e3 / (20 * nbatch):.2f} us per step (traced build)")
t = trace.view(600, 8)[:128].cpu().double()
names = ["start", "ptr+sync", "copy done", "scan+gather", "attention", "barrier", "head mma+sync", "end"]
print(f"streams={nstreams}: phase durations in SM cycles (median / p90 / max over 128 blocks of the last launches)")
for k in range(1, 8):
    d = t[:, k] - t[:, k - 1]
    print(f"  {names[k-1]:>14s} -> {names[k]:14s} {d.median():9.0f} {d.quantile(0.9):9.0f} {d.max():9.0f}")
d = t[:, 7] - t[:, 0]
print(f"  {'block lifetime':>32s} {d.median():9.0f} {d.quantile(0.9):9.0f} {d.max():9.0f}   ({d.median()/1965:.2f} us at 1965 MHz)")
