"""Developer tool: per-role cycle accounting of gemm_tf32x3_kernel (needs a -DQOT_TC_TRACE build of csrc/gemm_tc.cu
linked into scripts/libqot_b200_tctrace.so; QOT_B200_LIB=scripts/libqot_b200_tctrace.so python scripts/trace_gemm_tc.py [M K Nc])."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import ops, _lib
M, K, Nc = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 256, 1024)
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev); W = torch.randn(Nc, K, device=dev)
ops.gemm_tf32x3(A, W); torch.cuda.synchronize()
buf = torch.zeros(296 * 8, dtype=torch.int64, device=dev)
L = _lib.lib()
L.qot_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
assert L.qot_debug_set_tc_trace(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.gemm_tf32x3(A, W); e1.record(); torch.cuda.synchronize()
t = buf.view(296, 8).double()
used = t.sum(1) > 0
t = t[used]
kb = (M + 127) // 128 * ((Nc + 127) // 128) * (K // 32) / t.shape[0]
print(f"{e0.elapsed_time(e1) * 1e3:.1f} us incl. pre-pass; {t.shape[0]} CTAs, {kb:.0f} k-blocks per CTA")
names = ["producer: wait empty", "producer: issue copies", "mma: wait chain_free", "mma: wait full", "mma: issue + commit",
         "drain (warp 2): wait chain_full", "drain: tcgen05.ld + add + arrive", "drain: epilogue stores"]
for i, n in enumerate(names):
    print(f"{n:36s} mean {t[:, i].mean():10.0f} cycles/CTA = {t[:, i].mean() / kb:7.0f} per k-block  (min {t[:, i].min():.0f} max {t[:, i].max():.0f})")
