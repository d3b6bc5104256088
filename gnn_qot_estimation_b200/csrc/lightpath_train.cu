// LightpathGNN, general (training-capable) kernels over a destination-sorted CSR:
//   GATConv forward/backward (lightpath_training/models.py:13,30),
//   BatchNorm batch statistics + backward with a row-sparse upstream gradient
//   (models.py:14,31; A.5), ordered LUT compaction (models.py:35-40) and the MLP head
//   forward/backward (models.py:17-22,43).
//
// GAT rows use the reduced algebra  out_i[h,:] = W_h z_i^h,  z_i^h = sum_j alpha_ij^h x_j,
// s_j^h = x_j.(W_h^T att_src^h): messages are 5 floats.  Four lanes (one per head) own
// a destination row; its in-edges are consumed in edge order, so results are
// deterministic.  Weight gradients are fixed-order two-stage reductions.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int kF = QOT_GAT_IN, kHeads = QOT_GAT_HEADS, kC = QOT_GAT_C, kHC = QOT_GAT_HC, kHid = 32;
constexpr int kHeadStride = kC * kF + 1;   // 161: pads W so the 4 head-lanes hit distinct banks
constexpr int kGatRows = 64;               // destination rows per 256-thread block

struct GatSmem {
  float W[kHeads * kHeadStride];           // W[h][c][f]
  float As[kF][kHeads], Ad[kF][kHeads];
  float bias[kHC];
};

__device__ __forceinline__ void gat_load_params(GatSmem& s, const float* __restrict__ lin_w,
                                                const float* __restrict__ att_src,
                                                const float* __restrict__ att_dst,
                                                const float* __restrict__ bias) {
  for (int i = threadIdx.x; i < kHC * kF; i += blockDim.x) {
    const int c = i / kF, f = i % kF;
    s.W[(c / kC) * kHeadStride + (c % kC) * kF + f] = lin_w[i];
  }
  if (bias)
    for (int i = threadIdx.x; i < kHC; i += blockDim.x) s.bias[i] = bias[i];
  if (threadIdx.x < kF * kHeads) {
    const int f = threadIdx.x / kHeads, h = threadIdx.x % kHeads;
    float as = 0.f, ad = 0.f;
    for (int c = 0; c < kC; ++c) {
      const float w = lin_w[(h * kC + c) * kF + f];
      as = fmaf(w, att_src[h * kC + c], as);
      ad = fmaf(w, att_dst[h * kC + c], ad);
    }
    s.As[f][h] = as;
    s.Ad[f][h] = ad;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
gat_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ src, int64_t N, const float* __restrict__ lin_w,
               const float* __restrict__ att_src, const float* __restrict__ att_dst,
               const float* __restrict__ bias, float* __restrict__ hout, float* __restrict__ zout,
               float* __restrict__ smax, float* __restrict__ sden) {
  __shared__ GatSmem s;
  gat_load_params(s, lin_w, att_src, att_dst, bias);
  const int h = threadIdx.x & 3;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGatRows + (threadIdx.x >> 2);
  if (i >= N) return;
  float As[kF], xi[kF];
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    As[f] = s.As[f][h];
    xi[f] = x[i * kF + f];
  }
  float d = 0.f;
#pragma unroll
  for (int f = 0; f < kF; ++f) d = fmaf(xi[f], s.Ad[f][h], d);
  float m = -INFINITY, den = 0.f, acc[kF] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  for (int32_t p = beg; p < end; ++p) {
    const int64_t j = src[p];
    float xj[kF];
#pragma unroll
    for (int f = 0; f < kF; ++f) xj[f] = x[j * kF + f];
    float a = d;
#pragma unroll
    for (int f = 0; f < kF; ++f) a = fmaf(xj[f], As[f], a);
    a = a > 0.f ? a : 0.2f * a;
    const float mn = fmaxf(m, a);
    const float sc = expf(m - mn), pe = expf(a - mn);
    den = fmaf(den, sc, pe);
#pragma unroll
    for (int f = 0; f < kF; ++f) acc[f] = fmaf(acc[f], sc, pe * xj[f]);
    m = mn;
  }
  den += 1e-16f;
  float z[kF];
#pragma unroll
  for (int f = 0; f < kF; ++f) z[f] = acc[f] / den;
  if (zout) {
#pragma unroll
    for (int f = 0; f < kF; ++f) zout[i * (kHeads * kF) + h * kF + f] = z[f];
  }
  if (smax) {
    smax[i * kHeads + h] = m;
    sden[i * kHeads + h] = den;
  }
  const float* Wh = s.W + h * kHeadStride;
  float* orow = hout + i * kHC + h * kC;
#pragma unroll
  for (int c4 = 0; c4 < kC; c4 += 4) {
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c4 + k;
      float v = s.bias[h * kC + c];
#pragma unroll
      for (int f = 0; f < kF; ++f) v = fmaf(Wh[c * kF + f], z[f], v);
      o[k] = v;
    }
    *reinterpret_cast<float4*>(orow + c4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Backward, stage 1: per-row attention gradients folded into d(A_src), d(A_dst)
// block partials ([blocks][2][5][4]).
__global__ void __launch_bounds__(256)
gat_bwd_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ src, int64_t N, const float* __restrict__ lin_w,
                    const float* __restrict__ att_src, const float* __restrict__ att_dst,
                    const float* __restrict__ z, const float* __restrict__ smax,
                    const float* __restrict__ sden, const float* __restrict__ dh,
                    float* __restrict__ part) {
  __shared__ GatSmem s;
  __shared__ float red[kGatRows][kHeads][2 * kF + 1];
  gat_load_params(s, lin_w, att_src, att_dst, nullptr);
  const int h = threadIdx.x & 3, rl = threadIdx.x >> 2;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGatRows + rl;
  float dAs[kF] = {0.f, 0.f, 0.f, 0.f, 0.f}, dAd[kF] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < N) {
    float As[kF], xi[kF], u[kF] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int f = 0; f < kF; ++f) {
      As[f] = s.As[f][h];
      xi[f] = x[i * kF + f];
    }
    const float* Wh = s.W + h * kHeadStride;
    const float* drow = dh + i * kHC + h * kC;
#pragma unroll
    for (int c4 = 0; c4 < kC; c4 += 4) {
      const float4 g = *reinterpret_cast<const float4*>(drow + c4);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int f = 0; f < kF; ++f) u[f] = fmaf(Wh[(c4 + k) * kF + f], gv[k], u[f]);
    }
    float d = 0.f, t = 0.f;
#pragma unroll
    for (int f = 0; f < kF; ++f) {
      d = fmaf(xi[f], s.Ad[f][h], d);
      t = fmaf(u[f], z[i * (kHeads * kF) + h * kF + f], t);
    }
    const float m = smax[i * kHeads + h], den = sden[i * kHeads + h];
    float dd = 0.f;
    const int32_t beg = rowptr[i], end = rowptr[i + 1];
    for (int32_t p = beg; p < end; ++p) {
      const int64_t j = src[p];
      float xj[kF];
#pragma unroll
      for (int f = 0; f < kF; ++f) xj[f] = x[j * kF + f];
      float pre = d, dal = 0.f;
#pragma unroll
      for (int f = 0; f < kF; ++f) {
        pre = fmaf(xj[f], As[f], pre);
        dal = fmaf(u[f], xj[f], dal);
      }
      const float a = pre > 0.f ? pre : 0.2f * pre;
      const float alpha = expf(a - m) / den;
      const float de = alpha * (dal - t) * (pre > 0.f ? 1.f : 0.2f);
      dd += de;
#pragma unroll
      for (int f = 0; f < kF; ++f) dAs[f] = fmaf(de, xj[f], dAs[f]);
    }
#pragma unroll
    for (int f = 0; f < kF; ++f) dAd[f] = dd * xi[f];
  }
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    red[rl][h][f] = dAs[f];
    red[rl][h][kF + f] = dAd[f];
  }
  __syncthreads();
  if (threadIdx.x < kHeads * 2 * kF) {
    const int hh = threadIdx.x / (2 * kF), k = threadIdx.x % (2 * kF);
    float acc = 0.f;
    for (int r = 0; r < kGatRows; ++r) acc += red[r][hh][k];
    // layout [which(2)][f][h]
    part[static_cast<int64_t>(blockIdx.x) * 40 + (k / kF) * 20 + (k % kF) * kHeads + hh] = acc;
  }
}

// Backward, stage 2 (one block): fixed-order sum of the block partials, then chain
// through A = W_h^T att into d_lin_w, d_att_src, d_att_dst.
__global__ void __launch_bounds__(256)
gat_bwd_finish_kernel(const float* __restrict__ part, int64_t nblocks, const float* __restrict__ wz,
                      const float* __restrict__ lin_w, const float* __restrict__ att_src,
                      const float* __restrict__ att_dst, float* __restrict__ d_lin_w,
                      float* __restrict__ d_att_src, float* __restrict__ d_att_dst) {
  __shared__ float dA[40];      // [which][f][h]
  __shared__ float red[256];
  // 40 values x up to 6 row-lanes each, fixed strides -> deterministic
  const int v = threadIdx.x % 40, lane = threadIdx.x / 40;   // lanes 0..5 (240 threads)
  float acc = 0.f;
  if (threadIdx.x < 240)
    for (int64_t b = lane; b < nblocks; b += 6) acc += part[b * 40 + v];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < 40) {
    float t = 0.f;
    for (int l = 0; l < 6; ++l) t += red[l * 40 + threadIdx.x];
    dA[threadIdx.x] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHC * kF; i += blockDim.x) {
    const int c = i / kF, f = i % kF, h = c / kC;
    // aggregation part: (dh^T z)[c][h*5+f]
    d_lin_w[i] = wz[c * (kHeads * kF) + h * kF + f] + att_src[c] * dA[f * kHeads + h] +
                 att_dst[c] * dA[20 + f * kHeads + h];
  }
  for (int c = threadIdx.x; c < kHC; c += blockDim.x) {
    const int h = c / kC;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int f = 0; f < kF; ++f) {
      s1 = fmaf(lin_w[c * kF + f], dA[f * kHeads + h], s1);
      s2 = fmaf(lin_w[c * kF + f], dA[20 + f * kHeads + h], s2);
    }
    d_att_src[c] = s1;
    d_att_dst[c] = s2;
  }
}

// ---------------------------------------------------------------------------
// BatchNorm statistics: per-chunk (mean, M2) then Chan's pairwise combination in a
// fixed order (double accumulators in the tiny second stage).
// ---------------------------------------------------------------------------
constexpr int kBnRows = 256;

__global__ void __launch_bounds__(256)
bn_stats_stage1_kernel(const float* __restrict__ h, int64_t N, int64_t C, float* __restrict__ pmean,
                       float* __restrict__ pm2) {
  __shared__ float s[8][33];
  __shared__ float smean[32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 32 + cx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kBnRows, r1 = min(N, r0 + kBnRows);
  float acc = 0.f;
  if (c < C)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += h[r * C + c];
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][cx];
    smean[cx] = t / static_cast<float>(r1 - r0);
  }
  __syncthreads();
  const float mu = smean[cx];
  acc = 0.f;
  if (c < C)
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      const float dlt = h[r * C + c] - mu;
      acc = fmaf(dlt, dlt, acc);
    }
  __syncthreads();
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][cx];
    pmean[static_cast<int64_t>(blockIdx.y) * C + c] = mu;
    pm2[static_cast<int64_t>(blockIdx.y) * C + c] = t;
  }
}

__global__ void bn_stats_stage2_kernel(const float* __restrict__ pmean, const float* __restrict__ pm2,
                                       int64_t N, int64_t C, int64_t chunks, float* __restrict__ mean,
                                       float* __restrict__ var, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float momentum) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= C) return;
  double n_a = 0.0, mu = 0.0, m2 = 0.0;
  for (int64_t k = 0; k < chunks; ++k) {
    const double n_b = static_cast<double>(min(static_cast<int64_t>(kBnRows), N - k * kBnRows));
    const double mb = pmean[k * C + c], m2b = pm2[k * C + c];
    const double tot = n_a + n_b, dlt = mb - mu;
    mu += dlt * n_b / tot;
    m2 += m2b + dlt * dlt * n_a * n_b / tot;
    n_a = tot;
  }
  const double v = N > 0 ? m2 / static_cast<double>(N) : 0.0;
  mean[c] = static_cast<float>(mu);
  var[c] = static_cast<float>(v);
  if (running_mean) {
    const double unb = N > 1 ? m2 / static_cast<double>(N - 1) : v;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unb);
  }
}

// ---------------------------------------------------------------------------
// LUT compaction
// ---------------------------------------------------------------------------
__global__ void lut_flag_kernel(const float* __restrict__ x, int64_t N, int64_t F, int col,
                                int32_t* __restrict__ flag) {
  const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (n < N) flag[n] = (x[n * F + col] == 1.0f) ? 1 : 0;
}

__global__ void lut_scatter_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                   const int64_t* __restrict__ batch, int64_t N,
                                   int32_t* __restrict__ lut_node, int64_t* __restrict__ lut_batch,
                                   int32_t* __restrict__ n_lut) {
  const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (n == 0) n_lut[0] = pos[N];
  if (n < N && flag[n]) {
    lut_node[pos[n]] = static_cast<int32_t>(n);
    if (lut_batch) lut_batch[pos[n]] = batch[n];
  }
}

// ---------------------------------------------------------------------------
// LUT head: one warp per LUT row.
// ---------------------------------------------------------------------------
constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32)
lut_head_fwd_kernel(const float* __restrict__ h, const int32_t* __restrict__ lut_node, int64_t L,
                    const float* __restrict__ bn_mean, const float* __restrict__ bn_var, float eps,
                    const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                    const float* __restrict__ W1, const float* __restrict__ b1,
                    const float* __restrict__ W2, const float* __restrict__ b2,
                    const float* __restrict__ hmask, float* __restrict__ y, float* __restrict__ hid,
                    float* __restrict__ out) {
  __shared__ float s_w1t[kHC * kHid];
  __shared__ float s_y[kHeadWarps][kHC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kHC * kHid; i += blockDim.x) {
    const int c = i / kHid, o = i % kHid;
    s_w1t[i] = W1[o * kHC + c];
  }
  float scale[4], shift[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = k * 32 + lane;
    scale[k] = bn_w[c] / sqrtf(bn_var[c] + eps);
    shift[k] = bn_b[c] - bn_mean[c] * scale[k];
  }
  const float bb1 = b1[lane];
  float w2[QOT_OUT];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) w2[k] = W2[k * kHid + lane];
  const float bb2 = lane < QOT_OUT ? b2[lane] : 0.f;
  __syncthreads();
  for (int64_t l = static_cast<int64_t>(blockIdx.x) * kHeadWarps + warp; l < L;
       l += static_cast<int64_t>(gridDim.x) * kHeadWarps) {
    const int64_t node = lut_node[l];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = k * 32 + lane;
      const float v = fmaxf(fmaf(h[node * kHC + c], scale[k], shift[k]), 0.f);
      s_y[warp][c] = v;
      if (y) y[l * kHC + c] = v;
    }
    __syncwarp();
    float hv = bb1;
#pragma unroll 8
    for (int c = 0; c < kHC; c += 4) {
      const float4 yv = *reinterpret_cast<const float4*>(&s_y[warp][c]);
      hv = fmaf(s_w1t[(c + 0) * kHid + lane], yv.x, hv);
      hv = fmaf(s_w1t[(c + 1) * kHid + lane], yv.y, hv);
      hv = fmaf(s_w1t[(c + 2) * kHid + lane], yv.z, hv);
      hv = fmaf(s_w1t[(c + 3) * kHid + lane], yv.w, hv);
    }
    if (hid) hid[l * kHid + lane] = hv;
    float act = hv > 0.f ? hv : 0.01f * hv;
    if (hmask) act *= hmask[l * kHid + lane];
    const float o0 = warp_sum(act * w2[0]), o1 = warp_sum(act * w2[1]), o2 = warp_sum(act * w2[2]);
    if (lane < QOT_OUT) out[l * QOT_OUT + lane] = (lane == 0 ? o0 : lane == 1 ? o1 : o2) + bb2;
    __syncwarp();
  }
}

// Backward rows: dhid [L,32] (in place of `dhid_out`), act [L,32], dy [L,128].
__global__ void __launch_bounds__(kHeadWarps * 32)
lut_head_bwd_rows_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                         const float* __restrict__ hid, const float* __restrict__ hmask, int64_t L,
                         const float* __restrict__ W1, const float* __restrict__ W2,
                         float* __restrict__ dy, float* __restrict__ dhid_out, float* __restrict__ act_out) {
  __shared__ float s_w1[kHid * (kHC + 1)];      // [o][c], padded: lane c reads column c
  __shared__ float s_dh[kHeadWarps][kHid];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kHC * kHid; i += blockDim.x) s_w1[(i / kHC) * (kHC + 1) + (i % kHC)] = W1[i];
  float w2[QOT_OUT];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) w2[k] = W2[k * kHid + lane];
  __syncthreads();
  for (int64_t l = static_cast<int64_t>(blockIdx.x) * kHeadWarps + warp; l < L;
       l += static_cast<int64_t>(gridDim.x) * kHeadWarps) {
    const float hv = hid[l * kHid + lane];
    const float mk = hmask ? hmask[l * kHid + lane] : 1.f;
    float dact = 0.f;
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) dact = fmaf(w2[k], dout[l * QOT_OUT + k], dact);
    const float dh = dact * mk * (hv > 0.f ? 1.f : 0.01f);
    dhid_out[l * kHid + lane] = dh;
    act_out[l * kHid + lane] = (hv > 0.f ? hv : 0.01f * hv) * mk;
    s_dh[warp][lane] = dh;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = k * 32 + lane;
      float g = 0.f;
#pragma unroll 8
      for (int o = 0; o < kHid; ++o) g = fmaf(s_w1[o * (kHC + 1) + c], s_dh[warp][o], g);
      dy[l * kHC + c] = y[l * kHC + c] > 0.f ? g : 0.f;
    }
    __syncwarp();
  }
}

// d_bn_w / d_bn_b partials over LUT rows (chunks of 64 rows x 128 channels).
__global__ void __launch_bounds__(128)
bn_bwd_param_stage1_kernel(const float* __restrict__ h, const float* __restrict__ mean,
                           const float* __restrict__ var, float eps, const float* __restrict__ dy,
                           const int32_t* __restrict__ lut_node, int64_t L, int64_t C,
                           float* __restrict__ part) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
  if (c >= C) return;
  const int64_t l0 = static_cast<int64_t>(blockIdx.y) * 64, l1 = min(L, l0 + 64);
  const float mu = mean[c], inv = 1.f / sqrtf(var[c] + eps);
  float gw = 0.f, gb = 0.f;
  for (int64_t l = l0; l < l1; ++l) {
    const float g = dy[l * C + c];
    gw = fmaf(g, (h[static_cast<int64_t>(lut_node[l]) * C + c] - mu) * inv, gw);
    gb += g;
  }
  part[(static_cast<int64_t>(blockIdx.y) * 2 + 0) * C + c] = gw;
  part[(static_cast<int64_t>(blockIdx.y) * 2 + 1) * C + c] = gb;
}

__global__ void bn_bwd_param_stage2_kernel(const float* __restrict__ part, int64_t chunks, int64_t C,
                                           float* __restrict__ d_w, float* __restrict__ d_b) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= C) return;
  float gw = 0.f, gb = 0.f;
  for (int64_t k = 0; k < chunks; ++k) {
    gw += part[(k * 2 + 0) * C + c];
    gb += part[(k * 2 + 1) * C + c];
  }
  d_w[c] = gw;
  d_b[c] = gb;
}

// dh for every row: the batch-statistics terms (training) or zero (eval statistics).
__global__ void bn_bwd_dense_kernel(const float* __restrict__ h, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps,
                                    const float* __restrict__ bn_w, const float* __restrict__ d_w,
                                    const float* __restrict__ d_b, int64_t N, int64_t C, int batch_stats,
                                    float* __restrict__ dh) {
  const int64_t idx4 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // float4 index
  const int64_t total4 = N * C / 4;
  if (idx4 >= total4) return;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (batch_stats) {
    const int64_t c0 = (idx4 * 4) % C;
    const float4 hv = *reinterpret_cast<const float4*>(h + idx4 * 4);
    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
    float r[4];
    const float invN = 1.f / static_cast<float>(N);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t c = c0 + k;
      const float inv = 1.f / sqrtf(var[c] + eps);
      const float xhat = (hh[k] - mean[c]) * inv;
      r[k] = -bn_w[c] * inv * invN * (d_b[c] + xhat * d_w[c]);
    }
    o = make_float4(r[0], r[1], r[2], r[3]);
  }
  *reinterpret_cast<float4*>(dh + idx4 * 4) = o;
}

__global__ void bn_bwd_lut_rows_kernel(const float* __restrict__ var, float eps,
                                       const float* __restrict__ bn_w, const float* __restrict__ dy,
                                       const int32_t* __restrict__ lut_node, int64_t L, int64_t C,
                                       float* __restrict__ dh) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= L * C) return;
  const int64_t l = idx / C, c = idx % C;
  dh[static_cast<int64_t>(lut_node[l]) * C + c] += bn_w[c] / sqrtf(var[c] + eps) * dy[idx];
}

// ---------------------------------------------------------------------------
// Stand-alone BatchNorm over all rows (PyG BatchNorm used as its own layer,
// lightpath_training/models.py:31 when the reference's models.py is kept and only its layer
// imports are swapped): apply, and the dense backward.
// ---------------------------------------------------------------------------
__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t total4, int64_t C,
                                const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ y) {
  const int64_t i4 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i4 >= total4) return;
  const int64_t c0 = (i4 * 4) % C;
  const float4 v = *reinterpret_cast<const float4*>(x + i4 * 4);
  const float in[4] = {v.x, v.y, v.z, v.w};
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t c = c0 + k;
    o[k] = (in[k] - mean[c]) / sqrtf(var[c] + eps) * w[c] + b[c];
  }
  *reinterpret_cast<float4*>(y + i4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// partials of sum(dy * xhat) and sum(dy) per channel over chunks of kBnRows rows
__global__ void __launch_bounds__(256)
bn_bwd_dense_stage1_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                           const float* __restrict__ var, float eps, const float* __restrict__ dy,
                           int64_t N, int64_t C, float* __restrict__ part) {
  __shared__ float s_w[8][33], s_b[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 32 + cx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kBnRows, r1 = min(N, r0 + kBnRows);
  float gw = 0.f, gb = 0.f;
  if (c < C) {
    const float mu = mean[c], inv = 1.f / sqrtf(var[c] + eps);
#pragma unroll 4
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      const float g = dy[r * C + c];
      gw = fmaf(g, (x[r * C + c] - mu) * inv, gw);
      gb += g;
    }
  }
  s_w[ry][cx] = gw;
  s_b[ry][cx] = gb;
  __syncthreads();
  if (ry == 0 && c < C) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      tw += s_w[k][cx];
      tb += s_b[k][cx];
    }
    part[(static_cast<int64_t>(blockIdx.y) * 2 + 0) * C + c] = tw;
    part[(static_cast<int64_t>(blockIdx.y) * 2 + 1) * C + c] = tb;
  }
}

__global__ void bn_bwd_dense_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                          const float* __restrict__ var, float eps,
                                          const float* __restrict__ w, const float* __restrict__ dy,
                                          const float* __restrict__ d_w, const float* __restrict__ d_b,
                                          int64_t N, int64_t C, int batch_stats, float* __restrict__ dx) {
  const int64_t i4 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i4 >= N * C / 4) return;
  const int64_t c0 = (i4 * 4) % C;
  const float4 xv = *reinterpret_cast<const float4*>(x + i4 * 4);
  const float4 gv = *reinterpret_cast<const float4*>(dy + i4 * 4);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
  const float invN = 1.f / static_cast<float>(N);
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t c = c0 + k;
    const float inv = 1.f / sqrtf(var[c] + eps);
    float g = gs[k];
    if (batch_stats) g -= invN * (d_b[c] + (xs[k] - mean[c]) * inv * d_w[c]);
    o[k] = w[c] * inv * g;
  }
  *reinterpret_cast<float4*>(dx + i4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

}  // namespace qot

using namespace qot;

// ---------------------------------------------------------------------------
extern "C" int qot_gat_fwd(const float* x, const int32_t* rowptr, const int32_t* src, int64_t N,
                           const float* lin_w, const float* att_src, const float* att_dst,
                           const float* conv_bias, float* h, float* z, float* smax, float* sden,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0, "qot_gat_fwd: negative size");
  if (N == 0) return QOT_OK;
  QOT_REQUIRE(x && rowptr && src && lin_w && att_src && att_dst && conv_bias && h, "qot_gat_fwd: null argument");
  QOT_REQUIRE((smax == nullptr) == (sden == nullptr), "qot_gat_fwd: smax and sden go together");
  gat_fwd_kernel<<<static_cast<unsigned>(cdiv(N, kGatRows)), 256, 0, stream>>>(
      x, rowptr, src, N, lin_w, att_src, att_dst, conv_bias, h, z, smax, sden);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_gat_bwd_workspace_bytes(int64_t N) {
  if (N < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(N, 1), kGatRows)) * 40 * 4) +
         align_up(kHC * kHeads * kF * 4) + qot_wgrad_workspace_bytes(N, kHC, kHeads * kF) +
         qot_colsum_workspace_bytes(N, kHC) + 256;
}

extern "C" int qot_gat_bwd(const float* x, const int32_t* rowptr, const int32_t* src, int64_t N,
                           const float* lin_w, const float* att_src, const float* att_dst,
                           const float* z, const float* smax, const float* sden, const float* dh,
                           float* d_lin_w, float* d_att_src, float* d_att_dst, float* d_bias,
                           void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0, "qot_gat_bwd: negative size");
  QOT_REQUIRE(d_lin_w && d_att_src && d_att_dst && d_bias && lin_w && att_src && att_dst, "qot_gat_bwd: null argument");
  QOT_REQUIRE(N == 0 || (x && rowptr && src && z && smax && sden && dh), "qot_gat_bwd: null input");
  QOT_REQUIRE(ws && ws_bytes >= qot_gat_bwd_workspace_bytes(N), "qot_gat_bwd: workspace too small");
  const int64_t nblocks = cdiv(N, kGatRows);
  Carver c(ws);
  float* part = c.take<float>(std::max<int64_t>(nblocks, 1) * 40);
  float* wz = c.take<float>(kHC * kHeads * kF);
  const size_t wg_bytes = qot_wgrad_workspace_bytes(N, kHC, kHeads * kF);
  void* wg_ws = c.take<char>(wg_bytes);
  const size_t cs_bytes = qot_colsum_workspace_bytes(N, kHC);
  void* cs_ws = c.take<char>(cs_bytes);
  if (N > 0) {
    gat_bwd_rows_kernel<<<static_cast<unsigned>(nblocks), 256, 0, stream>>>(
        x, rowptr, src, N, lin_w, att_src, att_dst, z, smax, sden, dh, part);
    QOT_LAUNCH_CHECK();
  }
  int rc = qot_wgrad(dh, kHC, z, kHeads * kF, N, kHC, kHeads * kF, wz, kHeads * kF, wg_ws, wg_bytes, stream_);
  if (rc) return rc;
  if (N == 0) QOT_CUDA(cudaMemsetAsync(wz, 0, kHC * kHeads * kF * 4, stream));
  rc = qot_colsum(dh, kHC, N, kHC, d_bias, cs_ws, cs_bytes, stream_);
  if (rc) return rc;
  gat_bwd_finish_kernel<<<1, 256, 0, stream>>>(part, nblocks, wz, lin_w, att_src, att_dst, d_lin_w, d_att_src, d_att_dst);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_bn_stats_workspace_bytes(int64_t N, int64_t C) {
  if (N < 0 || C < 0) return 0;
  return 2 * align_up(static_cast<size_t>(cdiv(std::max<int64_t>(N, 1), kBnRows)) * C * 4) + 256;
}

extern "C" int qot_bn_stats(const float* h, int64_t N, int64_t C, float* mean, float* var,
                            float* running_mean, float* running_var, float momentum,
                            void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N > 0 && C > 0, "qot_bn_stats: need at least one row and one channel");
  QOT_REQUIRE(h && mean && var, "qot_bn_stats: null argument");
  QOT_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "qot_bn_stats: running stats go together");
  QOT_REQUIRE(ws && ws_bytes >= qot_bn_stats_workspace_bytes(N, C), "qot_bn_stats: workspace too small");
  const int64_t chunks = cdiv(N, kBnRows);
  QOT_REQUIRE(chunks <= 65535, "qot_bn_stats: too many rows for one launch");
  Carver c(ws);
  float* pmean = c.take<float>(chunks * C);
  float* pm2 = c.take<float>(chunks * C);
  dim3 grid(static_cast<unsigned>(cdiv(C, 32)), static_cast<unsigned>(chunks));
  bn_stats_stage1_kernel<<<grid, 256, 0, stream>>>(h, N, C, pmean, pm2);
  QOT_LAUNCH_CHECK();
  bn_stats_stage2_kernel<<<static_cast<unsigned>(cdiv(C, 128)), 128, 0, stream>>>(
      pmean, pm2, N, C, chunks, mean, var, running_mean, running_var, momentum);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_lut_select_workspace_bytes(int64_t N) {
  if (N < 0) return 0;
  return 2 * align_up(static_cast<size_t>(N + 1) * 4) + scan_workspace_bytes(N) + 256;
}

extern "C" int qot_lut_select(const float* x, int64_t N, int64_t F, int32_t is_lut_index,
                              const int64_t* batch, int32_t* lut_node, int64_t* lut_batch,
                              int32_t* n_lut, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && F > 0 && is_lut_index >= 0 && is_lut_index < F, "qot_lut_select: bad shape");
  QOT_REQUIRE(N < (1ll << 31) - 1, "qot_lut_select: N exceeds int32 range");
  QOT_REQUIRE(n_lut && (N == 0 || (x && lut_node)), "qot_lut_select: null argument");
  QOT_REQUIRE(!lut_batch || batch || N == 0, "qot_lut_select: lut_batch needs batch");
  QOT_REQUIRE(ws && ws_bytes >= qot_lut_select_workspace_bytes(N), "qot_lut_select: workspace too small");
  Carver c(ws);
  int32_t* flag = c.take<int32_t>(N + 1);
  int32_t* pos = c.take<int32_t>(N + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(N));
  if (N > 0) {
    lut_flag_kernel<<<static_cast<unsigned>(cdiv(N, 256)), 256, 0, stream>>>(x, N, F, is_lut_index, flag);
    QOT_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i32(flag, 0, pos, N, scan_ws, stream);
  if (rc) return rc;
  lut_scatter_kernel<<<static_cast<unsigned>(cdiv(std::max<int64_t>(N, 1), 256)), 256, 0, stream>>>(
      flag, pos, batch, N, lut_node, lut_batch, n_lut);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_lut_head_fwd(const float* h, const int32_t* lut_node, int64_t L,
                                const float* bn_mean, const float* bn_var, float bn_eps,
                                const float* bn_w, const float* bn_b, const float* W1, const float* b1,
                                const float* W2, const float* b2, const float* hmask, float* y,
                                float* hid, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(L >= 0, "qot_lut_head_fwd: negative size");
  if (L == 0) return QOT_OK;
  QOT_REQUIRE(h && lut_node && bn_mean && bn_var && bn_w && bn_b && W1 && b1 && W2 && b2 && out,
              "qot_lut_head_fwd: null argument");
  const int64_t blocks = std::min<int64_t>(cdiv(L, kHeadWarps), static_cast<int64_t>(kNumSMs) * 4);
  lut_head_fwd_kernel<<<static_cast<unsigned>(blocks), kHeadWarps * 32, 0, stream>>>(
      h, lut_node, L, bn_mean, bn_var, bn_eps, bn_w, bn_b, W1, b1, W2, b2, hmask, y, hid, out);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_lut_head_bwd_workspace_bytes(int64_t L) {
  if (L < 0) return 0;
  const int64_t l = std::max<int64_t>(L, 1);
  return 2 * align_up(static_cast<size_t>(l) * kHid * 4) + qot_wgrad_workspace_bytes(l, kHid, kHC) +
         qot_wgrad_workspace_bytes(l, QOT_OUT, kHid) + qot_colsum_workspace_bytes(l, kHid) + 256;
}

extern "C" int qot_lut_head_bwd(const float* dout, const float* y, const float* hid, const float* hmask,
                                int64_t L, const float* W1, const float* W2, float* dy, float* dW1,
                                float* db1, float* dW2, float* db2, void* ws, size_t ws_bytes,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(L > 0, "qot_lut_head_bwd: needs at least one LUT row");
  QOT_REQUIRE(dout && y && hid && W1 && W2 && dy && dW1 && db1 && dW2 && db2, "qot_lut_head_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_lut_head_bwd_workspace_bytes(L), "qot_lut_head_bwd: workspace too small");
  Carver c(ws);
  float* dhid = c.take<float>(L * kHid);
  float* act = c.take<float>(L * kHid);
  const size_t w1b = qot_wgrad_workspace_bytes(L, kHid, kHC), w2b = qot_wgrad_workspace_bytes(L, QOT_OUT, kHid);
  const size_t csb = qot_colsum_workspace_bytes(L, kHid);
  void* w1ws = c.take<char>(w1b);
  void* w2ws = c.take<char>(w2b);
  void* csws = c.take<char>(csb);
  const int64_t blocks = std::min<int64_t>(cdiv(L, kHeadWarps), static_cast<int64_t>(kNumSMs) * 4);
  lut_head_bwd_rows_kernel<<<static_cast<unsigned>(blocks), kHeadWarps * 32, 0, stream>>>(
      dout, y, hid, hmask, L, W1, W2, dy, dhid, act);
  QOT_LAUNCH_CHECK();
  int rc;
  if ((rc = qot_wgrad(dhid, kHid, y, kHC, L, kHid, kHC, dW1, kHC, w1ws, w1b, stream_))) return rc;
  if ((rc = qot_colsum(dhid, kHid, L, kHid, db1, csws, csb, stream_))) return rc;
  if ((rc = qot_wgrad(dout, QOT_OUT, act, kHid, L, QOT_OUT, kHid, dW2, kHid, w2ws, w2b, stream_))) return rc;
  if ((rc = qot_colsum(dout, QOT_OUT, L, QOT_OUT, db2, csws, csb, stream_))) return rc;
  return QOT_OK;
}

extern "C" size_t qot_bn_bwd_workspace_bytes(int64_t N, int64_t L, int64_t C) {
  if (N < 0 || L < 0 || C < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(L, 1), 64)) * 2 * C * 4) + 256;
}

extern "C" int qot_bn_bwd_sparse(const float* h, const float* mean, const float* var, float eps,
                                 const float* bn_w, const float* dy, const int32_t* lut_node,
                                 int64_t L, int64_t N, int64_t C, int batch_stats, float* dh,
                                 float* d_bn_w, float* d_bn_b, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N > 0 && L > 0 && C > 0 && C % 4 == 0, "qot_bn_bwd_sparse: bad shape");
  QOT_REQUIRE(h && mean && var && bn_w && dy && lut_node && dh && d_bn_w && d_bn_b, "qot_bn_bwd_sparse: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_bn_bwd_workspace_bytes(N, L, C), "qot_bn_bwd_sparse: workspace too small");
  const int64_t chunks = cdiv(L, 64);
  QOT_REQUIRE(chunks <= 65535, "qot_bn_bwd_sparse: too many LUT rows for one launch");
  float* part = static_cast<float*>(ws);
  dim3 g1(static_cast<unsigned>(cdiv(C, 128)), static_cast<unsigned>(chunks));
  bn_bwd_param_stage1_kernel<<<g1, 128, 0, stream>>>(h, mean, var, eps, dy, lut_node, L, C, part);
  QOT_LAUNCH_CHECK();
  bn_bwd_param_stage2_kernel<<<static_cast<unsigned>(cdiv(C, 128)), 128, 0, stream>>>(part, chunks, C, d_bn_w, d_bn_b);
  QOT_LAUNCH_CHECK();
  bn_bwd_dense_kernel<<<static_cast<unsigned>(cdiv(N * C / 4, 256)), 256, 0, stream>>>(
      h, mean, var, eps, bn_w, d_bn_w, d_bn_b, N, C, batch_stats, dh);
  QOT_LAUNCH_CHECK();
  bn_bwd_lut_rows_kernel<<<static_cast<unsigned>(cdiv(L * C, 256)), 256, 0, stream>>>(var, eps, bn_w, dy, lut_node, L, C, dh);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_bn_apply(const float* x, int64_t N, int64_t C, const float* mean, const float* var,
                            float eps, const float* w, const float* b, float* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && C > 0 && C % 4 == 0, "qot_bn_apply: C must be a positive multiple of 4");
  if (N == 0) return QOT_OK;
  QOT_REQUIRE(x && mean && var && w && b && y, "qot_bn_apply: null argument");
  const int64_t total4 = N * C / 4;
  bn_apply_kernel<<<static_cast<unsigned>(cdiv(total4, 256)), 256, 0, stream>>>(x, total4, C, mean, var, eps, w, b, y);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_bn_bwd_dense_workspace_bytes(int64_t N, int64_t C) {
  if (N < 0 || C < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(N, 1), kBnRows)) * 2 * C * 4) + 256;
}

extern "C" int qot_bn_bwd_dense(const float* x, const float* mean, const float* var, float eps,
                                const float* w, const float* dy, int64_t N, int64_t C, int batch_stats,
                                float* dx, float* d_w, float* d_b, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N > 0 && C > 0 && C % 4 == 0, "qot_bn_bwd_dense: bad shape");
  QOT_REQUIRE(x && mean && var && w && dy && dx && d_w && d_b, "qot_bn_bwd_dense: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_bn_bwd_dense_workspace_bytes(N, C), "qot_bn_bwd_dense: workspace too small");
  const int64_t chunks = cdiv(N, kBnRows);
  QOT_REQUIRE(chunks <= 65535, "qot_bn_bwd_dense: too many rows for one launch");
  float* part = static_cast<float*>(ws);
  dim3 g1(static_cast<unsigned>(cdiv(C, 32)), static_cast<unsigned>(chunks));
  bn_bwd_dense_stage1_kernel<<<g1, 256, 0, stream>>>(x, mean, var, eps, dy, N, C, part);
  QOT_LAUNCH_CHECK();
  bn_bwd_param_stage2_kernel<<<static_cast<unsigned>(cdiv(C, 128)), 128, 0, stream>>>(part, chunks, C, d_w, d_b);
  QOT_LAUNCH_CHECK();
  bn_bwd_dense_apply_kernel<<<static_cast<unsigned>(cdiv(N * C / 4, 256)), 256, 0, stream>>>(
      x, mean, var, eps, w, dy, d_w, d_b, N, C, batch_stats, dx);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
