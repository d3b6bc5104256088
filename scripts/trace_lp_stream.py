"""Developer tool: per-role cycle accounting of lp_stream_kernel (needs a -DQOT_ST_TRACE build:
scripts/build_variant.sh trace -DQOT_ST_TRACE; QOT_B200_LIB=scripts/libqot_b200_trace2.so python scripts/trace_lp_stream.py)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic, _lib
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", map_location="cpu", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
store = synthetic.lightpath_store(nb * 4096, seed=1, device=dev); store.verify_layout()
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(nb)]
split = len(sys.argv) > 2 and sys.argv[2] == "1"
plan = m.stream_plan(bs, split_head=split)
m.forward_stream(plan); torch.cuda.synchronize()
buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
L = _lib.lib()
L.qot_debug_set_st_trace.argtypes = [ctypes.c_void_p]
assert L.qot_debug_set_st_trace(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.forward_stream(plan); e1.record(); torch.cuda.synchronize()
t = buf.view(148, 8).double()
tiles = nb * 256 / 148
print(f"launch {e0.elapsed_time(e1) * 1e3:.1f} us for {nb} batches ({e0.elapsed_time(e1) * 1e3 / nb:.2f} us/batch incl. head kernel)")
names = ["producer wait-empty", "producer issue", "consumer wait-full (sum over warps)", "consumer work (sum over warps)",
         "consumer wait free z buffer (sum)", "epilogue wait staged group (4 warps)", "epilogue wait MMAs (4 warps)", "epilogue total (4 warps)"]
for i, n in enumerate(names):
    print(f"{n:40s} mean {t[:, i].mean():12.0f} cycles/CTA  = {t[:, i].mean() / tiles:9.0f} per tile   (min {t[:, i].min():.0f} max {t[:, i].max():.0f})")
