"""TEMP experiment: attn-only launches on S streams and head-only launches on S other streams, no
dependencies between them, in one graph."""
import sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic, _lib
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "mix"
nbatch = 96
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
store = synthetic.lightpath_store(nbatch * 4096, seed=1, device=dev)
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(nbatch)]
outs = [m.forward_device(b) for b in bs]
outs2 = [m.forward_device(b) for b in bs]
torch.cuda.synchronize()
L = _lib.lib()
side = torch.cuda.Stream()
sa = [torch.cuda.Stream() for _ in range(S)]
sh = [torch.cuda.Stream() for _ in range(S)]


def run():
    cur = torch.cuda.current_stream()
    for b in sa + sh:
        b.wait_stream(cur)
    for i in range(nbatch):
        if mode in ("mix", "attn"):
            L.qot_debug_lp_flags(2)      # skip head
            with torch.cuda.stream(sa[i % S]):
                m.forward_device(bs[i], outs[i])
        if mode in ("mix", "head"):
            L.qot_debug_lp_flags(1)      # skip attn
            with torch.cuda.stream(sh[i % S]):
                m.forward_device(bs[i], outs2[i])
    L.qot_debug_lp_flags(0)
    for b in sa + sh:
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
print(f"mode={mode} streams={S}+{S}: {ev0.elapsed_time(ev1) * 1e3 / (50 * nbatch):.3f} us per step")
