"""Developer tool: per-phase cycle counts of topo_fused_bwd_kernel (needs scripts/build_trace.sh)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import TopologicalGNN, synthetic, _lib
dev = torch.device("cuda:0")
m = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(dev)
store = synthetic.nsfnet_store(1024, seed=0, device=dev)
b = store.collate(range(1024))
L = _lib.lib()
L.qot_debug_set_tf_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(2048 * 24, dtype=torch.int64, device=dev)
for _ in range(3):
    m.zero_grad(); torch.nn.SmoothL1Loss()(m(b), b.y.view(-1, 3)).backward()
torch.cuda.synchronize()
assert L.qot_debug_set_tf_trace(trace.data_ptr()) == 0
m.zero_grad(); torch.nn.SmoothL1Loss()(m(b), b.y.view(-1, 3)).backward()
torch.cuda.synchronize()
t = trace.view(2048, 24)[:1024, :17].cpu().double()
names = ["fwd: inputs", "fwd: csr", "fwd: proj+hid", "fwd: attention", "fwd: T + O2", "fwd: pool+head (+dout)", "bwd: head", "bwd: dO2, Wroot, dH1",
         "bwd: dhid, W1", "bwd: dT", "bwd: dP, dH1, dO1", "bwd: skip", "bwd: attention dst", "bwd: attention src, We", "bwd: Wq/k/v, dX", "bwd: gemb"]
print("phase: median cycles over 1024 blocks")
for k, nme in enumerate(names):
    d = t[:, k + 1] - t[:, k]
    print(f"  {nme:28s} {d.median():9.0f}")
print(f"  {'total':28s} {(t[:, 16] - t[:, 0]).median():9.0f}")
