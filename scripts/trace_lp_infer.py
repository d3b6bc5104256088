"""Developer tool: per-block phase timeline of lp_infer_kernel (needs a -DQOT_LP_TRACE build:
scripts/build_trace.sh, then QOT_B200_LIB=scripts/libqot_b200_trace.so python scripts/trace_lp_infer.py)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic, _lib
dev = torch.device("cuda:0")
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
nbatch = 40
store = synthetic.lightpath_store(nbatch * 4096, seed=1, device=dev)
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(nbatch)]
outs = [m.forward_device(b) for b in bs]
torch.cuda.synchronize()
L = _lib.lib()
L.qot_debug_set_lp_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(600 * 8, dtype=torch.int64, device=dev)
for i in range(20, 30):
    m.forward_device(bs[i], outs[i])
torch.cuda.synchronize()
assert L.qot_debug_set_lp_trace(trace.data_ptr()) == 0
m.forward_device(bs[35], outs[35])
torch.cuda.synchronize()
nblk = 4096 // 8
t = trace.view(600, 8)[:nblk].cpu().double()
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0   # us
names = ["start", "extents", "loaded+sync", "-", "-", "-", "-", "end"]
print("phase: min / median / max  (us since first block start)")
for k, nme in enumerate(names):
    if nme == "-":
        continue
    c = rel[:, k]
    print(f"{nme:14s} {c.min():8.2f} {c.median():8.2f} {c.max():8.2f}")
pairs = [(0, 1), (1, 2), (2, 7)]
print("phase durations (median / max us):")
for a, b in pairs:
    d = rel[:, b] - rel[:, a]
    print(f"  {names[a]:>14s} -> {names[b]:14s} {d.median():7.2f} {d.max():7.2f}")
print("block start spread by tile (first 8, last 8):", rel[:8, 0].tolist(), rel[-8:, 0].tolist())
print("block end   by tile (first 8, last 8):", rel[:8, 7].tolist(), rel[-8:, 7].tolist())
