// Fused SmoothL1 loss (forward value + gradient) and running regression metrics, so the training /
// evaluation loops need no per-batch host synchronisation:
//   reference  topological_training/train.py:114-120  criterion(out, y); loss.item(); .cpu() per batch for R^2
//              lightpath_training/train.py:130-136     likewise on out vs y[lut_batch]
// loss = mean over all elements of smooth_l1(pred - target; beta) (torch.nn.SmoothL1Loss default
// reduction), dpred = d loss / d pred.  The metric accumulator keeps, per output column k and in fp64,
//   [count, sum y, sum y^2, sum (y - pred)^2, sum loss elements]
// across batches on the device; R^2_k = 1 - SS_res / (sum y^2 - (sum y)^2 / n) is formed by the caller
// (sklearn.r2_score's definition) when it finally reads the 15 numbers.
// Deterministic: per-block partials in a fixed order, one block finishes.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int kLossRows = 256;   // rows per first-stage block

__global__ void __launch_bounds__(256)
smooth_l1_stage1_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                        const int64_t* __restrict__ target_rows, int64_t n, float beta, float inv_count,
                        float* __restrict__ dpred, double* __restrict__ part) {
  __shared__ double s[8][QOT_OUT][4];
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kLossRows + threadIdx.x;
  double acc[QOT_OUT][4];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0;
  if (r < n) {
    const int64_t tr = target_rows ? target_rows[r] : r;
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) {
      const float p = pred[r * QOT_OUT + k], y = target[tr * QOT_OUT + k];
      const float d = p - y, ad = fabsf(d);
      const float l = ad < beta ? 0.5f * d * d / beta : ad - 0.5f * beta;
      if (dpred) dpred[r * QOT_OUT + k] = (ad < beta ? d / beta : (d > 0.f ? 1.f : -1.f)) * inv_count;
      acc[k][0] = y;
      acc[k][1] = static_cast<double>(y) * y;
      const double dd = static_cast<double>(p) - static_cast<double>(y);
      acc[k][2] = dd * dd;
      acc[k][3] = l;
    }
  }
  // warp tree, then the 8 warps in order
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = acc[k][q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k][q] = v;
    }
  __syncthreads();
  if (threadIdx.x < QOT_OUT * 4) {
    const int k = threadIdx.x / 4, q = threadIdx.x % 4;
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s[w][k][q];
    part[static_cast<int64_t>(blockIdx.x) * (QOT_OUT * 4) + threadIdx.x] = t;
  }
}

__global__ void smooth_l1_stage2_kernel(const double* __restrict__ part, int64_t nblocks, int64_t n,
                                        float* __restrict__ loss, double* __restrict__ metrics) {
  __shared__ double tot[QOT_OUT * 4];
  if (threadIdx.x < QOT_OUT * 4) {
    double t = 0.0;
    for (int64_t b = 0; b < nblocks; ++b) t += part[b * (QOT_OUT * 4) + threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    for (int k = 0; k < QOT_OUT; ++k) l += tot[k * 4 + 3];
    if (loss) *loss = static_cast<float>(l / static_cast<double>(std::max<int64_t>(n, 1) * QOT_OUT));
  }
  if (metrics && threadIdx.x < QOT_OUT) {
    const int k = threadIdx.x;
    metrics[k * 5 + 0] += static_cast<double>(n);
    metrics[k * 5 + 1] += tot[k * 4 + 0];
    metrics[k * 5 + 2] += tot[k * 4 + 1];
    metrics[k * 5 + 3] += tot[k * 4 + 2];
    metrics[k * 5 + 4] += tot[k * 4 + 3];
  }
}

}  // namespace qot

using namespace qot;

extern "C" size_t qot_smooth_l1_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(n, 1), kLossRows)) * QOT_OUT * 4 * sizeof(double)) + 256;
}

// pred [n,3]; target [*,3] read at row target_rows[r] (NULL: row r) -- `y[lut_batch]` of
// lightpath_training/train.py:122 without materialising it; loss [1] (optional), dpred [n,3] (optional),
// metrics [3,5] fp64 (optional, ACCUMULATED into; the caller zeroes it at epoch start).
extern "C" int qot_smooth_l1(const float* pred, const float* target, const int64_t* target_rows, int64_t n,
                             float beta, float* loss, float* dpred, double* metrics, void* ws,
                             size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(n > 0 && beta > 0.f, "qot_smooth_l1: needs at least one row and beta > 0");
  QOT_REQUIRE(pred && target, "qot_smooth_l1: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_smooth_l1_workspace_bytes(n) && (reinterpret_cast<uintptr_t>(ws) & 7) == 0,
              "qot_smooth_l1: workspace too small or misaligned");
  const int64_t nb = cdiv(n, kLossRows);
  double* part = static_cast<double*>(ws);
  smooth_l1_stage1_kernel<<<static_cast<unsigned>(nb), 256, 0, stream>>>(
      pred, target, target_rows, n, beta, 1.0f / static_cast<float>(n * QOT_OUT), dpred, part);
  QOT_LAUNCH_CHECK();
  smooth_l1_stage2_kernel<<<1, 32, 0, stream>>>(part, nb, n, loss, metrics);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
