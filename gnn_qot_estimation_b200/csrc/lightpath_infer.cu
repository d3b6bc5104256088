// LightpathGNN eval-mode forward, fully fused:
//   GATConv(5 -> 4x32, self loops replaced) -> BatchNorm(running stats) -> ReLU
//   -> LUT-row readout -> Linear(128,32) -> LeakyReLU -> Linear(32,3)
// (lightpath_training/models.py:26-45 under model.eval(); PyG semantics per
// SURVEY.md Appendix A.3/A.5).
//
// Graph-parallel: one warp walks one graph.  The node slab (20 B/node) and the
// destination row of edge_index (8 B/edge) are streamed once with coalesced loads;
// the source row is touched only for in-edges of a LUT node.  Only the rows the
// readout keeps are evaluated, in the algebraically reduced form
//   out_i[h,:] = W_h (sum_j alpha_ij^h x_j),   s_j^h = x_j . (W_h^T att_src^h)
// so a message is 5 floats, not 128.  Deterministic: edges of a row are consumed in
// edge order, reductions use fixed shuffle trees, no atomics on floats.
#include <algorithm>

#include "common.cuh"

namespace qot {

// prepared-parameter layout (floats)
constexpr int kF = QOT_GAT_IN;       // 5
constexpr int kHeads = QOT_GAT_HEADS;  // 4
constexpr int kC = QOT_GAT_C;        // 32
constexpr int kHC = QOT_GAT_HC;      // 128
constexpr int kHid = 32;             // mlp hidden
constexpr int kOffAsrc = 0;                       // [f][h] 20
constexpr int kOffAdst = kOffAsrc + kF * kHeads;  // [f][h] 20
constexpr int kOffWf = kOffAdst + kF * kHeads;    // [c][f] 640, BN scale folded in
constexpr int kOffShift = kOffWf + kHC * kF;      // [c] 128 (conv bias + BN folded)
constexpr int kOffW1t = kOffShift + kHC;          // [c][o] 4096
constexpr int kOffB1 = kOffW1t + kHC * kHid;      // [o] 32
constexpr int kOffW2 = kOffB1 + kHid;             // [k][o] 96
constexpr int kOffB2 = kOffW2 + QOT_OUT * kHid;   // [k] 3
constexpr int kPreparedFloats = ((kOffB2 + QOT_OUT + 3) / 4) * 4;

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
    for (int f = 0; f < kF; ++f) out[kOffWf + t * kF + f] = p.lin_w[t * kF + f] * scale;
    out[kOffShift + t] = (p.conv_bias[t] - p.bn_mean[t]) * scale + p.bn_b[t];
  }
  for (int i = t; i < kHC * kHid; i += blockDim.x) {
    const int c = i / kHid, o = i % kHid;
    out[kOffW1t + i] = p.mlp_w1[o * kHC + c];
  }
  if (t < kHid) out[kOffB1 + t] = p.mlp_b1[t];
  if (t < QOT_OUT * kHid) out[kOffW2 + t] = p.mlp_w2[t];
  if (t < QOT_OUT) out[kOffB2 + t] = p.mlp_b2[t];
}

// Phase A: LUT nodes per graph (also pulls the node slab through L2).
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

constexpr int kInferWarps = 8;

// Phase B: one warp per graph.
__global__ void __launch_bounds__(kInferWarps * 32)
lp_infer_kernel(const float* __restrict__ x, const int64_t* __restrict__ edge_index, int64_t E,
                const int64_t* __restrict__ gptr, const int64_t* __restrict__ eptr, int64_t N,
                int64_t B, const float* __restrict__ prep, int lut_col,
                const int32_t* __restrict__ lut_off, float* __restrict__ out,
                int64_t* __restrict__ lut_batch, int32_t* __restrict__ lut_node,
                int32_t* __restrict__ n_lut) {
  __shared__ float s_w1t[kHC * kHid];                 // 16 KB, [c][o]: lane o -> conflict-free
  __shared__ float s_z[kInferWarps][32];
  __shared__ float s_y[kInferWarps][kHC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kHC * kHid; i += blockDim.x) s_w1t[i] = prep[kOffW1t + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) n_lut[0] = lut_off[B];

  // lane role in the aggregation: head h, feature slot f (slots 5..7 idle)
  const int h = lane >> 3, f = lane & 7;
  float As[kF], Ad[kF];
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    As[k] = prep[kOffAsrc + k * kHeads + h];
    Ad[k] = prep[kOffAdst + k * kHeads + h];
  }
  // lane role in the projection: channels h2*32 + lane
  float Wf[kHeads][kF], shift[kHeads];
#pragma unroll
  for (int h2 = 0; h2 < kHeads; ++h2) {
    const int c = h2 * kC + lane;
#pragma unroll
    for (int k = 0; k < kF; ++k) Wf[h2][k] = prep[kOffWf + c * kF + k];
    shift[h2] = prep[kOffShift + c];
  }
  const float b1 = prep[kOffB1 + lane];
  float w2[QOT_OUT];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) w2[k] = prep[kOffW2 + k * kHid + lane];
  const float b2 = lane < QOT_OUT ? prep[kOffB2 + lane] : 0.f;
  __syncthreads();

  const int64_t* __restrict__ esrc = edge_index;
  const int64_t* __restrict__ edst = edge_index + E;
  const int64_t warp0 = blockIdx.x * static_cast<int64_t>(kInferWarps) + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kInferWarps;

  for (int64_t g = warp0; g < B; g += nwarps) {
    const int64_t n0 = gptr[g], n1 = gptr[g + 1];
    const int64_t e0 = eptr[g], e1 = eptr[g + 1];
    int32_t orow = lut_off[g];
    for (int64_t nb = n0; nb < n1; nb += 32) {
      const int64_t node = nb + lane;
      const bool is_lut = node < n1 && x[node * kF + lut_col] == 1.0f;
      unsigned mask = __ballot_sync(kFull, is_lut);
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int64_t i = nb + bit;
        // ---- attention over the in-edges of LUT node i (self loops replaced) ----
        float xi[kF];
#pragma unroll
        for (int k = 0; k < kF; ++k) xi[k] = x[i * kF + k];
        float d_i = 0.f, s_i = 0.f;
#pragma unroll
        for (int k = 0; k < kF; ++k) {
          d_i = fmaf(xi[k], Ad[k], d_i);
          s_i = fmaf(xi[k], As[k], s_i);
        }
        float m = -INFINITY, ssum = 0.f, acc = 0.f;
        for (int64_t eb = e0; eb < e1; eb += 32) {
          const int64_t e = eb + lane;
          const int64_t dd = (e < e1) ? edst[e] : -1;
          int64_t sj = i;
          if (dd == i) sj = esrc[e];
          const bool hit = (dd == i) && (sj != i) &&
                           (static_cast<uint64_t>(sj) < static_cast<uint64_t>(N));
          unsigned hm = __ballot_sync(kFull, hit);
          const int sj32 = static_cast<int>(sj - n0);   // graph-local: fits 32 bit
          while (hm) {
            const int l = __ffs(hm) - 1;
            hm &= hm - 1;
            const int64_t j = n0 + __shfl_sync(kFull, sj32, l);
            float xj[kF];
#pragma unroll
            for (int k = 0; k < kF; ++k) xj[k] = x[j * kF + k];
            float a = d_i;
#pragma unroll
            for (int k = 0; k < kF; ++k) a = fmaf(xj[k], As[k], a);
            a = a > 0.f ? a : 0.2f * a;
            const float mn = fmaxf(m, a);
            const float sc = expf(m - mn);      // exp(-inf) = 0 on the first edge
            const float pe = expf(a - mn);
            const float xv = (f == 0) ? xj[0] : (f == 1) ? xj[1] : (f == 2) ? xj[2]
                           : (f == 3) ? xj[3] : xj[4];
            ssum = fmaf(ssum, sc, pe);
            acc = fmaf(acc, sc, pe * xv);
            m = mn;
          }
        }
        {   // the appended self loop comes last, as in PyG's add_self_loops
          float a = s_i + d_i;
          a = a > 0.f ? a : 0.2f * a;
          const float mn = fmaxf(m, a);
          const float sc = expf(m - mn);
          const float pe = expf(a - mn);
          const float xv = (f == 0) ? xi[0] : (f == 1) ? xi[1] : (f == 2) ? xi[2]
                         : (f == 3) ? xi[3] : xi[4];
          ssum = fmaf(ssum, sc, pe);
          acc = fmaf(acc, sc, pe * xv);
        }
        s_z[warp][lane] = acc / (ssum + 1e-16f);   // z[h][f] at slot h*8+f
        __syncwarp();
        // ---- projection + folded BatchNorm + ReLU:  y[c], c = h2*32 + lane ----
#pragma unroll
        for (int h2 = 0; h2 < kHeads; ++h2) {
          float v = shift[h2];
#pragma unroll
          for (int k = 0; k < kF; ++k) v = fmaf(Wf[h2][k], s_z[warp][h2 * 8 + k], v);
          s_y[warp][h2 * kC + lane] = fmaxf(v, 0.f);
        }
        __syncwarp();
        // ---- MLP head: lane o owns hidden unit o ----
        float hid = b1;
#pragma unroll 8
        for (int c = 0; c < kHC; c += 4) {
          const float4 yv = *reinterpret_cast<const float4*>(&s_y[warp][c]);
          hid = fmaf(s_w1t[(c + 0) * kHid + lane], yv.x, hid);
          hid = fmaf(s_w1t[(c + 1) * kHid + lane], yv.y, hid);
          hid = fmaf(s_w1t[(c + 2) * kHid + lane], yv.z, hid);
          hid = fmaf(s_w1t[(c + 3) * kHid + lane], yv.w, hid);
        }
        hid = hid > 0.f ? hid : 0.01f * hid;
        float o0 = warp_sum(hid * w2[0]);
        float o1 = warp_sum(hid * w2[1]);
        float o2 = warp_sum(hid * w2[2]);
        const float ov = (lane == 0 ? o0 : lane == 1 ? o1 : o2) + b2;
        if (lane < QOT_OUT) out[static_cast<int64_t>(orow) * QOT_OUT + lane] = ov;
        if (lane == 0) {
          lut_batch[orow] = g;
          lut_node[orow] = static_cast<int32_t>(i);
        }
        ++orow;
        __syncwarp();
      }
    }
  }
}

}  // namespace qot

using namespace qot;

extern "C" size_t qot_lightpath_prepared_floats(void) { return kPreparedFloats; }

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

extern "C" size_t qot_lightpath_infer_workspace_bytes(int64_t N, int64_t B) {
  if (N < 0 || B < 0) return 0;
  return align_up(static_cast<size_t>(B + 1) * 4) * 2 + scan_workspace_bytes(B) + 256;
}

extern "C" int qot_lightpath_infer(const float* x, const int64_t* edge_index, int64_t E,
                                   const int64_t* gptr, const int64_t* eptr, int64_t N, int64_t B,
                                   const float* prepared, int32_t is_lut_index, float* out,
                                   int64_t* lut_batch, int32_t* lut_node, int32_t* n_lut,
                                   void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && B >= 0 && E >= 0, "qot_lightpath_infer: negative size");
  QOT_REQUIRE(N < (1ll << 31) - 1, "qot_lightpath_infer: N exceeds int32 range");
  QOT_REQUIRE(is_lut_index >= 0 && is_lut_index < kF, "qot_lightpath_infer: is_lut_index out of range");
  QOT_REQUIRE(gptr && eptr && prepared && n_lut, "qot_lightpath_infer: null argument");
  QOT_REQUIRE(N == 0 || (x && out && lut_batch && lut_node), "qot_lightpath_infer: null buffer");
  QOT_REQUIRE(E == 0 || edge_index, "qot_lightpath_infer: null edge_index");
  QOT_REQUIRE(ws && ws_bytes >= qot_lightpath_infer_workspace_bytes(N, B),
              "qot_lightpath_infer: workspace too small");
  Carver c(ws);
  int32_t* cnt = c.take<int32_t>(B + 1);
  int32_t* lut_off = c.take<int32_t>(B + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(B));
  if (B == 0) {
    QOT_CUDA(cudaMemsetAsync(n_lut, 0, 4, stream));
    return QOT_OK;
  }
  const int64_t blocks = std::min<int64_t>(cdiv(B, kInferWarps), static_cast<int64_t>(kNumSMs) * 8);
  lp_count_kernel<<<static_cast<unsigned>(blocks), kInferWarps * 32, 0, stream>>>(x, gptr, B, is_lut_index, cnt);
  QOT_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(cnt, 0, lut_off, B, scan_ws, stream);
  if (rc) return rc;
  lp_infer_kernel<<<static_cast<unsigned>(blocks), kInferWarps * 32, 0, stream>>>(
      x, edge_index, E, gptr, eptr, N, B, prepared, is_lut_index, lut_off, out, lut_batch, lut_node, n_lut);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
