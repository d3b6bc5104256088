"""Accuracy and time of the split-TF32 tensor-core GEMMs (csrc/gemm_tc.cu) against fp64, beside torch's fp32 GEMM, at
the BASELINE cfg 5 shapes.  GPU only.  profiles/r2_tc_chain_accuracy.md holds the sweep over the TMEM chain length
(TC_CHAIN in gemm_tc.cu; rebuild to change it) that this script produced.
usage: python scripts/tc_accuracy.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gnn_qot_estimation_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def err(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max())


A = torch.randn(10000, 256, generator=g).to(dev); W = torch.randn(1024, 256, generator=g).to(dev)
dy = torch.randn(80000, 256, generator=g).to(dev); x = torch.randn(80000, 256, generator=g).to(dev)
ref_f = A.double() @ W.double().t(); ref_w = dy.double().t() @ x.double()
print(f"torch fp32: fwd err {err(A @ W.t(), ref_f):.2e} ({timed(lambda: A @ W.t()):.1f} us)  "
      f"wgrad err {err(dy.t() @ x, ref_w):.2e} ({timed(lambda: dy.t() @ x):.1f} us)")
print(f"tf32x3    : fwd err {err(ops.gemm_tf32x3(A, W), ref_f):.2e} ({timed(lambda: ops.gemm_tf32x3(A, W)):.1f} us)  "
      f"wgrad err {err(ops.wgrad_tf32x3(dy, x), ref_w):.2e} ({timed(lambda: ops.wgrad_tf32x3(dy, x)):.1f} us)")
for (M, K, Nc) in ((10000, 256, 2560), (80000, 256, 256), (65536, 256, 1024)):
    A2 = torch.randn(M, K, generator=g).to(dev); W2 = torch.randn(Nc, K, generator=g).to(dev)
    print(f"[{M},{K}]x[{K},{Nc}]: tf32x3 {timed(lambda: ops.gemm_tf32x3(A2, W2)):.1f} us, torch fp32 {timed(lambda: A2 @ W2.t()):.1f} us")
