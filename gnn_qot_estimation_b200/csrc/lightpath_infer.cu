// LightpathGNN eval-mode forward, fully fused:
//   GATConv(5 -> 4x32, self loops replaced) -> BatchNorm(running stats) -> ReLU
//   -> LUT-row readout -> Linear(128,32) -> LeakyReLU -> Linear(32,3)
// (lightpath_training/models.py:26-45 under model.eval(); PyG semantics per
// SURVEY.md Appendix A.3/A.5).
//
// ONE launch per batch, graph-parallel: one warp owns one graph.  All of the graph's
// compulsory bytes are requested up front -- the node slab (20 B/node, staged in shared
// memory) and the destination row of edge_index (8 B/edge, kept in registers) -- with
// coalesced loads; the source row is touched only for in-edges of a LUT node.  Only the
// rows the readout keeps are evaluated, in the algebraically reduced form
//   out_i[h,:] = W_h (sum_j alpha_ij^h x_j),   s_j^h = x_j . (W_h^T att_src^h)
// so a message is 5 floats, not 128.  Output rows must appear in ascending node order
// across the whole batch (x[lut_mask]): the per-block LUT counts are chained with a
// decoupled look-back over a small self-cleaning state array, so no separate count/scan
// launches are needed.  Weights are read through L1 (ld.global.nc), not staged per block.
// Deterministic: edges of a row are consumed in edge order, reductions use fixed shuffle
// trees, no atomics on floats.
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
constexpr int kOffWf = kOffAdst + kF * kHeads;    // [h][f][c%32] 640, BN scale folded in
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
}

constexpr int kIW = 8;                    // warps (= graphs) per block
constexpr int kMaxN = 64;                 // fast path: nodes staged in shared memory
constexpr int kXF = kMaxN * kF;           // 320 floats per graph
constexpr int kXR = kXF / 32;             // 10 slab loads per lane
constexpr int kEC = 8;                    // fast path: 8 x 32 = 256 edges held in registers
constexpr unsigned long long kFlagAgg = 1ull << 32, kFlagInc = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// online-softmax update of one message (logit a, feature value xv of this lane's slot)
__device__ __forceinline__ void attn_update(float a, float xv, float& m, float& ssum, float& acc) {
  a = a > 0.f ? a : 0.2f * a;
  const float mn = fmaxf(m, a);
  const float sc = expf(m - mn);          // exp(-inf) = 0 on the first message
  const float pe = expf(a - mn);
  ssum = fmaf(ssum, sc, pe);
  acc = fmaf(acc, sc, pe * xv);
  m = mn;
}
__device__ __forceinline__ float pick5(const float (&v)[kF], int f) {
  return (f == 0) ? v[0] : (f == 1) ? v[1] : (f == 2) ? v[2] : (f == 3) ? v[3] : v[4];
}

// z (slot h*8+f of s_z) -> folded projection + BatchNorm + ReLU -> MLP head; lanes 0..2 return out[k]
__device__ __forceinline__ float lut_row_head(const float* __restrict__ prep, const float* s_z, float* s_y,
                                              int lane) {
  // y[c], c = h2*32 + lane
#pragma unroll
  for (int h2 = 0; h2 < kHeads; ++h2) {
    float v = __ldg(prep + kOffShift + h2 * kC + lane);
#pragma unroll
    for (int k = 0; k < kF; ++k) v = fmaf(__ldg(prep + kOffWf + (h2 * kF + k) * kC + lane), s_z[h2 * 8 + k], v);
    s_y[h2 * kC + lane] = fmaxf(v, 0.f);
  }
  __syncwarp();
  // hidden layer: lane = (cg, og) owns outputs 4og..4og+3 over channels [32cg, 32cg+32)
  const int cg = lane >> 3, og = lane & 7;
  float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* __restrict__ w1 = reinterpret_cast<const float4*>(prep + kOffW1t) + og;
#pragma unroll 8
  for (int cc = 0; cc < 32; ++cc) {
    const int c = cg * 32 + cc;
    const float4 w = __ldg(w1 + c * (kHid / 4));
    const float yv = s_y[c];
    h4.x = fmaf(w.x, yv, h4.x);
    h4.y = fmaf(w.y, yv, h4.y);
    h4.z = fmaf(w.z, yv, h4.z);
    h4.w = fmaf(w.w, yv, h4.w);
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    h4.x += __shfl_xor_sync(kFull, h4.x, o);
    h4.y += __shfl_xor_sync(kFull, h4.y, o);
    h4.z += __shfl_xor_sync(kFull, h4.z, o);
    h4.w += __shfl_xor_sync(kFull, h4.w, o);
  }
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(prep + kOffB1) + og);
  float4 act = make_float4(h4.x + b1.x, h4.y + b1.y, h4.z + b1.z, h4.w + b1.w);
  act.x = act.x > 0.f ? act.x : 0.01f * act.x;
  act.y = act.y > 0.f ? act.y : 0.01f * act.y;
  act.z = act.z > 0.f ? act.z : 0.01f * act.z;
  act.w = act.w > 0.f ? act.w : 0.01f * act.w;
  float o3[QOT_OUT];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(prep + kOffW2 + k * kHid) + og);
    float p = w.x * act.x + w.y * act.y + w.z * act.z + w.w * act.w;
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) p += __shfl_xor_sync(kFull, p, o);
    o3[k] = p;
  }
  const float b2 = lane < QOT_OUT ? __ldg(prep + kOffB2 + lane) : 0.f;
  return (lane == 0 ? o3[0] : lane == 1 ? o3[1] : o3[2]) + b2;
}

// Generic row evaluation straight from global memory (graphs beyond the fast-path caps).
__device__ float lut_row_global(const float* __restrict__ x, const int64_t* __restrict__ esrc,
                                const int64_t* __restrict__ edst, int64_t e0, int64_t e1, int64_t N,
                                int64_t i, const float (&As)[kF], const float (&Ad)[kF], int f,
                                const float* __restrict__ prep, float* s_z, float* s_y, int lane) {
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
    const bool hit = (dd == i) && (sj != i) && (static_cast<uint64_t>(sj) < static_cast<uint64_t>(N));
    unsigned hm = __ballot_sync(kFull, hit);
    while (hm) {
      const int l = __ffs(hm) - 1;
      hm &= hm - 1;
      const int64_t j = __shfl_sync(kFull, sj, l);
      float xj[kF];
#pragma unroll
      for (int k = 0; k < kF; ++k) xj[k] = x[j * kF + k];
      float a = d_i;
#pragma unroll
      for (int k = 0; k < kF; ++k) a = fmaf(xj[k], As[k], a);
      attn_update(a, pick5(xj, f), m, ssum, acc);
    }
  }
  attn_update(s_i + d_i, pick5(xi, f), m, ssum, acc);   // the appended self loop comes last
  __syncwarp();
  s_z[lane] = acc / (ssum + 1e-16f);
  __syncwarp();
  return lut_row_head(prep, s_z, s_y, lane);
}

__global__ void __launch_bounds__(kIW * 32, 4)
lp_infer_kernel(const float* __restrict__ x, const int64_t* __restrict__ edge_index, int64_t E,
                const int64_t* __restrict__ gptr, const int64_t* __restrict__ eptr, int64_t N,
                int64_t B, const float* __restrict__ prep, int lut_col,
                unsigned long long* __restrict__ st, float* __restrict__ out,
                int64_t* __restrict__ lut_batch, int32_t* __restrict__ lut_node,
                int32_t* __restrict__ n_lut) {
  __shared__ float s_x[kIW][kXF];
  __shared__ float s_y[kIW][kHC];
  __shared__ float s_z[kIW][32];
  __shared__ int s_cnt[kIW];
  __shared__ int s_red[kIW][3];
  __shared__ int s_base, s_flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = blockIdx.x, ntiles = gridDim.x;
  const int64_t g = tile * kIW + warp;
  const bool active = g < B;
  const int64_t* __restrict__ esrc = edge_index;
  const int64_t* __restrict__ edst = edge_index + E;

  // ---- (1) graph extents: 4 lanes fetch gptr[g], gptr[g+1], eptr[g], eptr[g+1]
  long long pv = 0;
  if (active && lane < 4) pv = (lane < 2) ? gptr[g + lane] : eptr[g + lane - 2];
  const int64_t n0 = __shfl_sync(kFull, pv, 0), n1 = __shfl_sync(kFull, pv, 1);
  const int64_t e0 = __shfl_sync(kFull, pv, 2), e1 = __shfl_sync(kFull, pv, 3);
  const bool fast = active && (n1 - n0) <= kMaxN && (e1 - e0) <= 32 * kEC && n1 >= n0 && e1 >= e0;
  const int n = fast ? static_cast<int>(n1 - n0) : 0;
  const int ne = fast ? static_cast<int>(e1 - e0) : 0;

  // ---- (2) every compulsory byte of the graph requested at once
  int dl[kEC];                                   // destination, graph-local (-1: not in this graph)
  if (fast) {
    const float* __restrict__ xg = x + n0 * kF;
    const int nf = n * kF;
    float xr[kXR];
#pragma unroll
    for (int k = 0; k < kXR; ++k) {
      const int idx = lane + 32 * k;
      xr[k] = idx < nf ? __ldg(xg + idx) : 0.f;
    }
    long long dq[kEC];
#pragma unroll
    for (int k = 0; k < kEC; ++k) {
      const int e = lane + 32 * k;
      dq[k] = e < ne ? edst[e0 + e] : -1;
    }
#pragma unroll
    for (int k = 0; k < kXR; ++k) s_x[warp][lane + 32 * k] = xr[k];
#pragma unroll
    for (int k = 0; k < kEC; ++k) {
      const long long dd = dq[k] - n0;
      dl[k] = (dq[k] >= 0 && dd >= 0 && dd < n) ? static_cast<int>(dd) : -1;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kEC; ++k) dl[k] = -1;
  }
  __syncwarp();

  // ---- (3) LUT flags of this graph, block aggregate published for the look-back
  unsigned m0 = 0, m1 = 0;
  int cnt = 0;
  if (fast) {
    m0 = __ballot_sync(kFull, lane < n && s_x[warp][lane * kF + lut_col] == 1.0f);
    m1 = __ballot_sync(kFull, lane + 32 < n && s_x[warp][(lane + 32) * kF + lut_col] == 1.0f);
    cnt = __popc(m0) + __popc(m1);
  } else if (active) {
    for (int64_t nb = n0; nb < n1; nb += 32) {
      const int64_t node = nb + lane;
      cnt += __popc(__ballot_sync(kFull, node < n1 && x[node * kF + lut_col] == 1.0f));
    }
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  int tot = 0, wexcl = 0;
#pragma unroll
  for (int w = 0; w < kIW; ++w) {
    const int c = s_cnt[w];
    if (w < warp) wexcl += c;
    tot += c;
  }
  if (threadIdx.x == 0)
    st_relaxed_u64(st + 1 + tile, (tile == 0 ? kFlagInc : kFlagAgg) | static_cast<unsigned int>(tot));

  // lane role in the aggregation: head h, feature slot f (slots 5..7 idle)
  const int h = lane >> 3, f = lane & 7;
  float As[kF], Ad[kF];
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    As[k] = __ldg(prep + kOffAsrc + k * kHeads + h);
    Ad[k] = __ldg(prep + kOffAdst + k * kHeads + h);
  }

  // fast-path row: attention over the in-edges of local node il, from registers + shared memory
  auto fast_row = [&](int il) -> float {
    const float* sx = s_x[warp];
    float xi[kF];
#pragma unroll
    for (int k = 0; k < kF; ++k) xi[k] = sx[il * kF + k];
    float d_i = 0.f, s_i = 0.f;
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      d_i = fmaf(xi[k], Ad[k], d_i);
      s_i = fmaf(xi[k], As[k], s_i);
    }
    // sources of the matching edges, all requests in flight together
    long long sq[kEC];
#pragma unroll
    for (int k = 0; k < kEC; ++k) sq[k] = (dl[k] == il) ? esrc[e0 + lane + 32 * k] : -1;
    float m = -INFINITY, ssum = 0.f, acc = 0.f;
#pragma unroll
    for (int k = 0; k < kEC; ++k) {
      if (32 * k >= ne) break;                                   // warp-uniform
      const long long sj = sq[k];
      const bool hit = (dl[k] == il) && (sj != n0 + il) &&
                       (static_cast<uint64_t>(sj) < static_cast<uint64_t>(N));
      unsigned hm = __ballot_sync(kFull, hit);
      const long long sloc = sj - n0;
      const int jl_mine = (sloc >= 0 && sloc < n) ? static_cast<int>(sloc) : -1;   // -1: outside the slab
      while (hm) {
        const int l = __ffs(hm) - 1;
        hm &= hm - 1;
        const int jl = __shfl_sync(kFull, jl_mine, l);
        float xj[kF];
        if (jl >= 0) {
#pragma unroll
          for (int k2 = 0; k2 < kF; ++k2) xj[k2] = sx[jl * kF + k2];
        } else {                                                 // cross-graph source: read it from global
          const long long j = __shfl_sync(kFull, sj, l);
#pragma unroll
          for (int k2 = 0; k2 < kF; ++k2) xj[k2] = x[j * kF + k2];
        }
        float a = d_i;
#pragma unroll
        for (int k2 = 0; k2 < kF; ++k2) a = fmaf(xj[k2], As[k2], a);
        attn_update(a, pick5(xj, f), m, ssum, acc);
      }
    }
    attn_update(s_i + d_i, pick5(xi, f), m, ssum, acc);          // appended self loop, last
    __syncwarp();
    s_z[warp][lane] = acc / (ssum + 1e-16f);
    __syncwarp();
    return lut_row_head(prep, s_z[warp], s_y[warp], lane);
  };
  auto nth_lut = [&](int r) -> int {                             // r-th LUT node of a fast graph
    unsigned a = m0;
    int base = 0;
    const int c0 = __popc(m0);
    if (r >= c0) { a = m1; r -= c0; base = 32; }
    for (int t = 0; t < r; ++t) a &= a - 1;
    return base + __ffs(a) - 1;
  };

  // ---- (4) first row of every fast graph before the output offset is known
  float ov0 = 0.f;
  int il0 = -1;
  if (fast && cnt > 0) {
    il0 = nth_lut(0);
    ov0 = fast_row(il0);
  }

  // ---- (5) exclusive prefix of the LUT counts of all preceding tiles (decoupled look-back,
  //          block-wide window of kIW*32 predecessors per round)
  if (tile == 0) {
    if (threadIdx.x == 0) s_base = 0;
  } else {
    int excl = 0;
    int64_t hi = tile - 1;
    while (true) {
      const int64_t idx = hi - threadIdx.x;
      unsigned long long wv = kFlagInc;                          // virtual tiles < 0: inclusive 0
      if (idx >= 0) wv = ld_relaxed_u64(st + 1 + idx);
      const unsigned flag = static_cast<unsigned>(wv >> 32);
      const int val = static_cast<int>(static_cast<unsigned>(wv));
      // per-warp: lanes are predecessors hi-32w-lane; closest inclusive = lowest lane with flag 2
      const unsigned inc_m = __ballot_sync(kFull, flag == 2);
      const unsigned emp_m = __ballot_sync(kFull, flag == 0);
      const int first_inc = inc_m ? __ffs(inc_m) - 1 : 32;
      const int first_emp = emp_m ? __ffs(emp_m) - 1 : 32;
      int part = (lane <= first_inc) ? val : 0;                  // tiles up to and incl. the inclusive one
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      if (lane == 0) {
        s_red[warp][0] = part;
        s_red[warp][1] = first_inc;
        s_red[warp][2] = first_emp;
      }
      __syncthreads();
      // combine the warps in predecessor order (warp 0 = closest tiles)
      int sum = 0;
      bool ready = true, found = false;
#pragma unroll
      for (int w = 0; w < kIW; ++w) {
        if (found || !ready) break;
        const int fi = s_red[w][1], fe = s_red[w][2];
        if (fe < fi) { ready = false; break; }                   // an unpublished tile before any inclusive one
        sum += s_red[w][0];
        if (fi < 32) found = true;
      }
      __syncthreads();
      if (!ready) continue;                                      // spin: re-read the window
      excl += sum;
      if (found) break;
      hi -= kIW * 32;
    }
    if (threadIdx.x == 0) s_base = excl;
  }
  __syncthreads();
  const int base = s_base;
  if (threadIdx.x == 0) {
    if (tile != 0) st_relaxed_u64(st + 1 + tile, kFlagInc | static_cast<unsigned int>(base + tot));
    if (tile == ntiles - 1) n_lut[0] = base + tot;
  }

  // ---- (6) outputs in ascending node order
  int64_t orow = static_cast<int64_t>(base) + wexcl;
  if (fast) {
    for (int r = 0; r < cnt; ++r) {
      const int il = (r == 0) ? il0 : nth_lut(r);
      const float ov = (r == 0) ? ov0 : fast_row(il);
      if (lane < QOT_OUT) out[orow * QOT_OUT + lane] = ov;
      if (lane == 0) {
        lut_batch[orow] = g;
        lut_node[orow] = static_cast<int32_t>(n0 + il);
      }
      ++orow;
    }
  } else if (active && cnt > 0) {
    for (int64_t nb = n0; nb < n1; nb += 32) {
      const int64_t node = nb + lane;
      unsigned mask = __ballot_sync(kFull, node < n1 && x[node * kF + lut_col] == 1.0f);
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int64_t i = nb + bit;
        const float ov = lut_row_global(x, esrc, edst, e0, e1, N, i, As, Ad, f, prep, s_z[warp], s_y[warp], lane);
        if (lane < QOT_OUT) out[orow * QOT_OUT + lane] = ov;
        if (lane == 0) {
          lut_batch[orow] = g;
          lut_node[orow] = static_cast<int32_t>(i);
        }
        ++orow;
      }
    }
  }

  // ---- (7) the last block through its look-back wipes the state for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long prev = atomicAdd(st, 1ull);
    s_flag = (prev == static_cast<unsigned long long>(ntiles - 1));
  }
  __syncthreads();
  if (s_flag) {
    for (int64_t t = threadIdx.x; t <= ntiles; t += blockDim.x) st[t] = 0ull;
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

extern "C" size_t qot_lightpath_infer_state_bytes(int64_t B) {
  if (B < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(B, 1), kIW) + 1) * sizeof(unsigned long long));
}

extern "C" int qot_lightpath_infer(const float* x, const int64_t* edge_index, int64_t E,
                                   const int64_t* gptr, const int64_t* eptr, int64_t N, int64_t B,
                                   const float* prepared, int32_t is_lut_index, float* out,
                                   int64_t* lut_batch, int32_t* lut_node, int32_t* n_lut,
                                   void* state, size_t state_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && B >= 0 && E >= 0, "qot_lightpath_infer: negative size");
  QOT_REQUIRE(N < (1ll << 31) - 1, "qot_lightpath_infer: N exceeds int32 range");
  QOT_REQUIRE(is_lut_index >= 0 && is_lut_index < kF, "qot_lightpath_infer: is_lut_index out of range");
  QOT_REQUIRE(gptr && eptr && prepared && n_lut, "qot_lightpath_infer: null argument");
  QOT_REQUIRE(N == 0 || (x && out && lut_batch && lut_node), "qot_lightpath_infer: null buffer");
  QOT_REQUIRE(E == 0 || edge_index, "qot_lightpath_infer: null edge_index");
  QOT_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 15) == 0, "qot_lightpath_infer: prepared must be 16-byte aligned");
  QOT_REQUIRE(state && state_bytes >= qot_lightpath_infer_state_bytes(B) &&
                  (reinterpret_cast<uintptr_t>(state) & 7) == 0,
              "qot_lightpath_infer: state buffer too small or misaligned");
  if (B == 0) {
    QOT_CUDA(cudaMemsetAsync(n_lut, 0, 4, stream));
    return QOT_OK;
  }
  const int64_t blocks = cdiv(B, kIW);
  QOT_REQUIRE(blocks < (1ll << 31) - 1, "qot_lightpath_infer: too many graphs for one launch");
  lp_infer_kernel<<<static_cast<unsigned>(blocks), kIW * 32, 0, stream>>>(
      x, edge_index, E, gptr, eptr, N, B, prepared, is_lut_index,
      static_cast<unsigned long long*>(state), out, lut_batch, lut_node, n_lut);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
