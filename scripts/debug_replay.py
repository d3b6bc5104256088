"""Debug: does a CUDA-graph replay of forward_device equal the eager call?"""
import sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import LightpathGNN, synthetic
dev = torch.device("cuda:0")
sd = torch.load("tests/golden/ckpt_lightpath_model_1.pt", weights_only=False)["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
store = synthetic.lightpath_store(3 * 4096, seed=1, device=dev)
bs = [store.collate(range(i * 4096, (i + 1) * 4096)) for i in range(3)]
outs = [m.forward_device(b) for b in bs]
torch.cuda.synchronize()
eager = [(o.out.clone(), o.lut_batch.clone(), int(o.n_lut.item())) for o in outs]
# eager determinism
for i, b in enumerate(bs):
    r = m.forward_device(b)
    n = int(r.n_lut.item())
    print("eager again", i, n, eager[i][2], torch.equal(r.out[:n], eager[i][0][:n]), torch.equal(r.lut_batch[:n], eager[i][1][:n]))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i, b in enumerate(bs):
            m.forward_device(b, outs[i])
    for o in outs:
        o.out.zero_(); o.n_lut.zero_()
    g.replay()
    torch.cuda.synchronize()
for i in range(3):
    n = int(outs[i].n_lut.item())
    d = (outs[i].out[:eager[i][2]] - eager[i][0][:eager[i][2]]).abs()
    bad = (d.max(dim=1).values > 0).nonzero().flatten()
    print("replay", i, "n", n, "eager n", eager[i][2], "maxdiff", float(d.max()), "bad rows", bad.numel(), bad[:10].tolist())
    ref, lb = m(bs[i])
    print("   module call equal eager:", torch.equal(ref, eager[i][0][:eager[i][2]]), ref.shape)
