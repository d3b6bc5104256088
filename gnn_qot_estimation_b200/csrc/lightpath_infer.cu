// LightpathGNN eval-mode forward -- parameter folding and the lut_ptr index array; the kernel itself is
// lp_stream_kernel (lightpath_stream.cu).  The forward, fully fused:
//   GATConv(5 -> 4x32, self loops replaced) -> BatchNorm(running stats) -> ReLU
//   -> LUT-row readout -> Linear(128,32) -> LeakyReLU -> Linear(32,3)
// (lightpath_training/models.py:26-45 under model.eval(); PyG semantics per
// SURVEY.md Appendix A.3/A.5).
//
// Graph-parallel: all of a graph's compulsory bytes are requested up front -- the node slab
// (20 B/node) and the destination row of edge_index (8 B/edge) -- the source row is touched only
// for in-edges of a LUT node.  Only the rows the readout keeps are evaluated, in the
// algebraically reduced form
//   out_i[h,:] = W_h (sum_j alpha_ij^h x_j),   s_j^h = x_j . (W_h^T att_src^h)
// so a message is 5 floats, not 128.  Output rows appear in ascending node order across the
// whole batch (x[lut_mask]) because every graph writes at lut_ptr[g], an index array the collate
// builds like ptr.  Deterministic: edges of a row are consumed in edge order, reductions use
// fixed trees, no atomics on floats.
// This file: lp_prepare_kernel (folded / pre-split / swizzled parameters), lp_count_kernel + scan (lut_ptr).
#include <algorithm>
#include <cstdlib>

#include "lightpath_common.cuh"

namespace qot {

__global__ void __launch_bounds__(256)
lp_prepare_kernel(qot_lightpath_params_t p, float* __restrict__ out) {
  const int t = threadIdx.x;
  // A_src[f][h] = sum_c W[h*32+c][f] * att_src[h][c]   (and A_dst likewise)
  if (t < kF * kHeads) {
    const int f = t / kHeads, h = t % kHeads;
    float as = 0.f, ad = 0.f;
    for (int c = 0; c < kC; ++c) {
      const float w = p.lin_w[(h * kC + c) * kF + f];
      as = fmaf(w, p.att_src[h * kC + c], as);
      ad = fmaf(w, p.att_dst[h * kC + c], ad);
    }
    out[kOffAsrc + f * kHeads + h] = as;
    out[kOffAdst + f * kHeads + h] = ad;
  }
  // BN(eval) folded: y = relu(scale*(W z + conv_bias - mean) + bn_b)
  if (t < kHC) {
    const float scale = p.bn_w[t] / sqrtf(p.bn_var[t] + p.bn_eps);
    for (int f = 0; f < kF; ++f)
      out[kOffWf + ((t / kC) * kF + f) * kC + (t % kC)] = p.lin_w[t * kF + f] * scale;
    out[kOffShift + t] = (p.conv_bias[t] - p.bn_mean[t]) * scale + p.bn_b[t];
  }
  for (int i = t; i < kHC * kHid; i += blockDim.x) {
    const int c = i / kHid, o = i % kHid;
    out[kOffW1t + i] = p.mlp_w1[o * kHC + c];
  }
  if (t < kHid) out[kOffB1 + t] = p.mlp_b1[t];
  if (t < QOT_OUT * kHid) out[kOffW2 + t] = p.mlp_w2[t];
  if (t < QOT_OUT) out[kOffB2 + t] = p.mlp_b2[t];
  __syncthreads();                                     // A_src / A_dst above are read back below
  for (int i = t; i < 8 * 6 * 4; i += blockDim.x) {
    const int sl = i / 24, m = (i / 4) % 6, hs = i % 4;
    const int fm = (sl & 1) ? (m + 3) % 6 : m;          // feature of slot m; 5 = the constant one
    out[kOffAsP + i] = fm == 5 ? 0.f : out[kOffAsrc + fm * kHeads + (hs ^ (sl >> 1))];
  }
  for (int i = t; i < 4 * kF * 4; i += blockDim.x) {
    const int hx = i / (kF * 4), f = (i / 4) % kF, hs = i % 4;
    out[kOffAdP + i] = out[kOffAdst + f * kHeads + (hs ^ hx)];
  }
  // B fragments (mma.m16n8k8 .col): lane holds B[k = t4][n = g8] and B[k = t4 + 4][n = g8]
  for (int i = t; i < 16 * 32; i += blockDim.x) {
    const int j = i >> 5, ln = i & 31, g8 = ln >> 2, t4 = ln & 3;
    const int c = 8 * j + g8;                              // output channel of the GAT layer
    const float scale = p.bn_w[c] / sqrtf(p.bn_var[c] + p.bn_eps);
    const float shift = (p.conv_bias[c] - p.bn_mean[c]) * scale + p.bn_b[c];
    // k slots 0..4: folded projection row, slot 5: the shift (its A entry is the constant 1), 6..7: zero
    const float b0 = p.lin_w[c * kF + t4] * scale;
    const float b1 = (t4 == 0) ? p.lin_w[c * kF + 4] * scale : (t4 == 1) ? shift : 0.f;
    reinterpret_cast<float4*>(out + kOffB1f)[i] = tf32_split2(b0, b1);
  }
  // mlp.0 weight in the swizzled K-major tcgen05 operand layout (lightpath_common.cuh: kOffB2sw)
  for (int i = t; i < 4 * 32 * 32; i += blockDim.x) {
    const int kb = i >> 10, n = (i >> 5) & 31, kk = i & 31;       // k = 32 kb + kk
    const int c = kk >> 2, j = kk & 3;
    const float w = p.mlp_w1[n * kHC + kb * 32 + kk];
    const float hi = __uint_as_float(tf32_rna(w));
    const int o = kb * 1024 + n * 32 + ((c ^ (n & 7)) << 2) + j;
    out[kOffB2sw + o] = hi;
    out[kOffB2sw + 4096 + o] = __uint_as_float(tf32_rna(w - hi));
  }
  for (int i = t; i < 16 * 4 * 32; i += blockDim.x) {
    const int j = i >> 7, q = (i >> 5) & 3, ln = i & 31, g8 = ln >> 2, t4 = ln & 3;
    // k slot t4 <-> channel 8j + 2*t4, slot t4 + 4 <-> channel 8j + 2*t4 + 1: the C fragment of the
    // first product is the A fragment of the second without any data movement
    const int c0 = 8 * j + 2 * t4, o = 8 * q + g8;
    reinterpret_cast<float4*>(out + kOffB2f)[i] = tf32_split2(p.mlp_w1[o * kHC + c0], p.mlp_w1[o * kHC + c0 + 1]);
  }
}

// lut_cnt[g] = number of nodes of graph g whose LUT flag is 1.0 (feeds the lut_ptr scan)
__global__ void __launch_bounds__(256)
lp_count_kernel(const float* __restrict__ x, const int64_t* __restrict__ gptr, int64_t B,
                int lut_col, int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t g = warp0; g < B; g += nwarps) {
    const int64_t n0 = gptr[g], n1 = gptr[g + 1];
    int c = 0;
    for (int64_t nb = n0; nb < n1; nb += 32) {
      const int64_t node = nb + lane;
      const bool f = node < n1 && x[node * kF + lut_col] == 1.0f;
      c += __popc(__ballot_sync(kFull, f));
    }
    if (lane == 0) cnt[g] = c;
  }
}
__global__ void widen_i32_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = in[i];
}

}  // namespace qot

using namespace qot;

extern "C" size_t qot_lightpath_prepared_floats(void) { return kPreparedFloats; }
extern "C" size_t qot_lightpath_infer_workspace_bytes(int64_t N) {
  return N > 0 ? align_up(static_cast<size_t>(N) * kHeads * kF * 4) : 256;
}

extern "C" int qot_lightpath_prepare(const qot_lightpath_params_t* p, float* prepared, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(p && prepared, "qot_lightpath_prepare: null argument");
  QOT_REQUIRE(p->lin_w && p->att_src && p->att_dst && p->conv_bias && p->bn_w && p->bn_b &&
                  p->bn_mean && p->bn_var && p->mlp_w1 && p->mlp_b1 && p->mlp_w2 && p->mlp_b2,
              "qot_lightpath_prepare: null parameter tensor");
  lp_prepare_kernel<<<1, 256, 0, stream>>>(*p, prepared);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_lightpath_lut_ptr_workspace_bytes(int64_t B) {
  if (B < 0) return 0;
  return 2 * align_up(static_cast<size_t>(B + 1) * 4) + scan_workspace_bytes(B) + 256;
}

extern "C" int qot_lightpath_lut_ptr(const float* x, const int64_t* gptr, int64_t N, int64_t B,
                                     int32_t is_lut_index, int64_t* lut_ptr, void* ws, size_t ws_bytes,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && B >= 0 && lut_ptr && gptr, "qot_lightpath_lut_ptr: bad argument");
  QOT_REQUIRE(N == 0 || x, "qot_lightpath_lut_ptr: null x");
  QOT_REQUIRE(is_lut_index >= 0 && is_lut_index < kF, "qot_lightpath_lut_ptr: is_lut_index out of range");
  QOT_REQUIRE(ws && ws_bytes >= qot_lightpath_lut_ptr_workspace_bytes(B), "qot_lightpath_lut_ptr: workspace too small");
  Carver c(ws);
  int32_t* cnt = c.take<int32_t>(B + 1);
  int32_t* off = c.take<int32_t>(B + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(B));
  if (B > 0) {
    const int64_t blocks = std::min<int64_t>(cdiv(B, 8), static_cast<int64_t>(kNumSMs) * 8);
    lp_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, gptr, B, is_lut_index, cnt);
    QOT_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i32(cnt, 0, off, B, scan_ws, stream);
  if (rc) return rc;
  widen_i32_kernel<<<static_cast<unsigned>(cdiv(B + 1, 256)), 256, 0, stream>>>(off, B + 1, lut_ptr);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

