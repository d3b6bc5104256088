"""GPU parity of the tensor-core projection (csrc/gemm_tc.cu: tcgen05.mma kind::tf32 on hi/lo-split
operands, TMEM accumulator) against fp64: the 3 x TF32 scheme must stay at fp32 accuracy (1e-5 bar),
which plain TF32 (~1e-3) would not."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


# tile-pair shapes of the cluster kernel: 1x1 and 3x1 tiles (odd x odd: pair along n with an off-the-edge partner),
# 8x2 / 33x8 / 79x20 (pair along n), 2x3 and 8x1 (n-tile count odd, m-tile count even: pair along m, W multicast),
# and a long reduction (128 k-blocks = 64 TMEM chains per tile over the four accumulators)
@pytest.mark.parametrize("M,Nc,K", [(128, 128, 32), (1000, 256, 64), (4097, 1024, 256), (10000, 2560, 256), (300, 68, 96),
                                    (256, 384, 64), (1024, 100, 128), (200, 256, 4096)])
@pytest.mark.parametrize("bias", [False, True])
def test_gemm_tf32x3_matches_fp64(cuda, M, Nc, K, bias):
    from gnn_qot_estimation_b200 import ops
    g = torch.Generator().manual_seed(M + Nc + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(Nc, K, generator=g) / K ** 0.5
    b = torch.randn(Nc, generator=g) if bias else None
    C = ops.gemm_tf32x3(A.to(cuda), W.to(cuda), b.to(cuda) if bias else None)
    ref = A.double() @ W.double().t() + (b.double() if bias else 0.0)
    assert int(ops._tc_status[C.device].item()) == 0
    assert rel_err(C, ref) <= 1e-5
    # entrywise: error relative to the magnitude of the terms summed (|A||W|), i.e. fp32-like
    scale = (A.abs().double() @ W.abs().double().t()).clamp_min(1e-30)
    assert float(((C.double().cpu() - ref).abs() / scale).max()) <= 2e-6


def test_gemm_tf32x3_gather_rows(cuda):
    from gnn_qot_estimation_b200 import ops
    g = torch.Generator().manual_seed(0)
    table = torch.randn(500, 128, generator=g)
    ids = torch.randint(0, 500, (3000,), generator=g)
    W = torch.randn(256, 128, generator=g) * 0.1
    C = ops.gemm_tf32x3(table.to(cuda), W.to(cuda), None, ids.to(cuda))
    ref = table[ids].double() @ W.double().t()
    assert rel_err(C, ref) <= 1e-5


def test_gemm_tf32x3_deterministic(cuda):
    from gnn_qot_estimation_b200 import ops
    A = torch.randn(2048, 256, device=cuda)
    W = torch.randn(512, 256, device=cuda)
    assert torch.equal(ops.gemm_tf32x3(A, W), ops.gemm_tf32x3(A, W))


@pytest.mark.parametrize("R,Mo,No", [(1000, 128, 128), (10000, 2560, 256), (4099, 1024, 256), (2048, 70, 100),
                                     (3000, 256, 128), (80000, 256, 256), (37, 128, 384)])
def test_wgrad_tf32x3_matches_fp64(cuda, R, Mo, No):
    """dW = dy^T x on the tensor cores (transposed split operands, split-K slices summed in fixed order)."""
    from gnn_qot_estimation_b200 import ops
    g = torch.Generator().manual_seed(R + Mo)
    dy = torch.randn(R, Mo, generator=g)
    x = torch.randn(R, No, generator=g)
    C = ops.wgrad_tf32x3(dy.to(cuda), x.to(cuda))
    ref = dy.double().t() @ x.double()
    assert int(ops._tc_status[C.device].item()) == 0
    assert rel_err(C, ref) <= 1e-5
    scale = (dy.abs().double().t() @ x.abs().double()).clamp_min(1e-30)
    assert float(((C.double().cpu() - ref).abs() / scale).max()) <= 2e-6
    assert torch.equal(C, ops.wgrad_tf32x3(dy.to(cuda), x.to(cuda)))


def test_wgrad_tf32x3_gather(cuda):
    from gnn_qot_estimation_b200 import ops
    g = torch.Generator().manual_seed(4)
    table = torch.randn(300, 128, generator=g)
    ids = torch.randint(0, 300, (5000,), generator=g)
    dy = torch.randn(5000, 192, generator=g)
    C = ops.wgrad_tf32x3(dy.to(cuda), table.to(cuda), ids.to(cuda))
    ref = dy.double().t() @ table[ids].double()
    assert rel_err(C, ref) <= 1e-5
