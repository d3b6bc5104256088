// TopologicalGNN edge-phase kernels over a destination-sorted CSR:
//   TransformerConv (topological_training/models.py:15-17,53),
//   NNConv(aggr=mean) in factorised form (models.py:20-30,57; A.2),
//   global_mean_pool + MLP head (models.py:33-38,61-63; A.4),
// and their backward kernels.  The node-wise projections are qot_gemm calls.
//
// Mapping: a group of LANES = H/4 lanes (capped at 32) owns one destination row; each
// lane holds VEC float4 slices of the H channels, so every gather of a neighbour row
// is a run of coalesced 128-bit loads.  In-edges are consumed in edge order with an
// online softmax; per-destination reductions are shuffles inside the group.  The
// backward scatters to sources through the transposed (source-sorted) CSR instead of
// atomics, so forward and backward are deterministic.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int K_ = QOT_EDGE_HID;   // 8
constexpr int D_ = QOT_EDGE_DIM;   // 4

struct EdgeMlp1 {                  // first layer of the edge MLP, register resident
  float w[K_][D_];
  float b[K_];
};
// block-wide load: global -> shared -> every thread's registers (call from all threads)
__device__ __forceinline__ EdgeMlp1 load_edge_mlp1(const float* __restrict__ W1, const float* __restrict__ b1) {
  __shared__ EdgeMlp1 s_w1;
  if (threadIdx.x < K_ * D_) s_w1.w[threadIdx.x / D_][threadIdx.x % D_] = W1[threadIdx.x];
  if (threadIdx.x < K_) s_w1.b[threadIdx.x] = b1[threadIdx.x];
  __syncthreads();
  return s_w1;
}

template <int LANES>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (LANES == 32) {
    return kFull;
  } else {
    const int lane = threadIdx.x & 31;
    return ((1u << LANES) - 1u) << (lane & ~(LANES - 1));
  }
}
template <int LANES>
__device__ __forceinline__ float gsum(float v, unsigned mask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 fma4(float s, const float4& a, const float4& c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float4 scale4(float s, const float4& a) {
  return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}
// derivative mask of leaky_relu evaluated on the ACTIVATED output (sign preserved)
__device__ __forceinline__ float4 lrelu_grad4(const float4& g, const float4& o, float slope) {
  return make_float4(o.x > 0.f ? g.x : g.x * slope, o.y > 0.f ? g.y : g.y * slope,
                     o.z > 0.f ? g.z : g.z * slope, o.w > 0.f ? g.w : g.w * slope);
}
__device__ __forceinline__ float4 lrelu4(const float4& v, float slope) {
  return make_float4(leaky(v.x, slope), leaky(v.y, slope), leaky(v.z, slope), leaky(v.w, slope));
}
// edge projection slice: e[c] = sum_d We[c][d] a[d] for the lane's 4 channels
__device__ __forceinline__ float4 edge_proj(const float4 (&w)[4], const float4& a) {
  return make_float4(dot4(w[0], a), dot4(w[1], a), dot4(w[2], a), dot4(w[3], a));
}

// ===========================================================================
// TransformerConv forward
// ===========================================================================
template <int LANES>
__device__ __forceinline__ float4 gshfl4(const float4& v, int srcl, unsigned mask) {
  return make_float4(__shfl_sync(mask, v.x, srcl, LANES), __shfl_sync(mask, v.y, srcl, LANES),
                     __shfl_sync(mask, v.z, srcl, LANES), __shfl_sync(mask, v.w, srcl, LANES));
}

// The in-edges of a row are consumed in chunks of LANES: lane l fetches the (src, eid, edge_attr) of
// edge c0+l with coalesced loads, the values are handed round by shuffles, and the k_j / v_j rows of
// edge t+1 are requested before edge t is reduced -- the row's gathers are no longer a chain of
// dependent round trips (index -> row -> index -> row ...).
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
tconv_fwd_kernel(const float* __restrict__ qkvs, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                 const float* __restrict__ edge_attr, const float* __restrict__ We, int64_t N,
                 float slope, float* __restrict__ out, float* __restrict__ logit,
                 float* __restrict__ rmax, float* __restrict__ rden) {
  constexpr int H = LANES * VEC * 4;
  const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / LANES;
  const int l = threadIdx.x % LANES;
  if (i >= N) return;
  const unsigned mask = group_mask<LANES>();
  const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(H));
  float4 q[VEC], acc[VEC], we[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = (v * LANES + l) * 4;
    q[v] = scale4(inv_sqrt, ld4(qkvs + i * 4 * H + c0));
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) we[v][k] = ld4(We + (c0 + k) * D_);
  }
  float m = -INFINITY, den = 0.f;
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  for (int32_t c0 = beg; c0 < end; c0 += LANES) {
    const int cnt = min(LANES, end - c0);
    int my_src = 0;
    float4 my_ea = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < cnt) {
      my_src = src[c0 + l];
      my_ea = ld4(edge_attr + static_cast<int64_t>(eid[c0 + l]) * D_);
    }
    float4 kj[VEC], vj[VEC];
    {
      const int64_t j = __shfl_sync(mask, my_src, 0, LANES);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = (v * LANES + l) * 4;
        kj[v] = ld4(qkvs + j * 4 * H + H + cc);
        vj[v] = ld4(qkvs + j * 4 * H + 2 * H + cc);
      }
    }
    for (int t = 0; t < cnt; ++t) {
      float4 kn[VEC], vn[VEC];
      if (t + 1 < cnt) {                                         // group-uniform: next edge's rows in flight
        const int64_t jn = __shfl_sync(mask, my_src, t + 1, LANES);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int cc = (v * LANES + l) * 4;
          kn[v] = ld4(qkvs + jn * 4 * H + H + cc);
          vn[v] = ld4(qkvs + jn * 4 * H + 2 * H + cc);
        }
      }
      const float4 ea = gshfl4<LANES>(my_ea, t, mask);
      float4 ve[VEC];
      float part = 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float4 e = edge_proj(we[v], ea);
        const float4 ke = add4(kj[v], e);
        ve[v] = add4(vj[v], e);
        part += dot4(q[v], ke);
      }
      const float a = gsum<LANES>(part, mask);
      if (logit && l == 0) logit[c0 + t] = a;
      const float mn = fmaxf(m, a);
      const float sc = expf(m - mn), pe = expf(a - mn);
      den = fmaf(den, sc, pe);
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = fma4(pe, ve[v], scale4(sc, acc[v]));
      m = mn;
      if (t + 1 < cnt) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          kj[v] = kn[v];
          vj[v] = vn[v];
        }
      }
    }
  }
  den += 1e-16f;
  const float inv = 1.0f / den;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = (v * LANES + l) * 4;
    const float4 skip = ld4(qkvs + i * 4 * H + 3 * H + c0);
    float4 o = fma4(inv, acc[v], skip);
    st4(out + i * H + c0, lrelu4(o, slope));
  }
  if (rmax && l == 0) {
    rmax[i] = m;
    rden[i] = den;
  }
}

// ===========================================================================
// TransformerConv backward, destination side.
//   g = dout * lrelu'(out);  dskip = g;  t = g . msg  (msg = preact - skip);
//   per edge: alpha, dalpha = g.(v_j+e), da = alpha (dalpha - t);
//   dq_i += da (k_j+e)/sqrt(H);  de = alpha g + da q_i/sqrt(H);  dWe += de a^T;
//   wa[eid] = alpha, wda[eid] = da/sqrt(H) feed the source-side kernel.
// ===========================================================================
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
tconv_bwd_dst_kernel(const float* __restrict__ qkvs, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                     const float* __restrict__ edge_attr, const float* __restrict__ We,
                     const float* __restrict__ out, const float* __restrict__ dout,
                     const float* __restrict__ logit, const float* __restrict__ rmax,
                     const float* __restrict__ rden, int64_t N, float slope,
                     float* __restrict__ dqkvs, float* __restrict__ wa, float* __restrict__ wda,
                     float* __restrict__ dwe_part) {
  constexpr int H = LANES * VEC * 4;
  constexpr int ROWS = 256 / LANES;
  extern __shared__ float s_red[];                 // [ROWS][H*4]
  const int rl = threadIdx.x / LANES, l = threadIdx.x % LANES;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * ROWS + rl;
  const unsigned mask = group_mask<LANES>();
  const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(H));
  float4 dwe[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int k = 0; k < 4; ++k) dwe[v][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < N) {
    float4 q[VEC], g[VEC], dq[VEC], we[VEC][4];
    float tpart = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c0 = (v * LANES + l) * 4;
      q[v] = ld4(qkvs + i * 4 * H + c0);
      const float4 o = ld4(out + i * H + c0);
      g[v] = lrelu_grad4(ld4(dout + i * H + c0), o, slope);
      st4(dqkvs + i * 4 * H + 3 * H + c0, g[v]);
      const float4 skip = ld4(qkvs + i * 4 * H + 3 * H + c0);
      const float inv_slope = 1.0f / slope;
      const float4 pre = make_float4(o.x > 0.f ? o.x : o.x * inv_slope, o.y > 0.f ? o.y : o.y * inv_slope,
                                     o.z > 0.f ? o.z : o.z * inv_slope, o.w > 0.f ? o.w : o.w * inv_slope);
      const float4 msg = make_float4(pre.x - skip.x, pre.y - skip.y, pre.z - skip.z, pre.w - skip.w);
      tpart += dot4(g[v], msg);
      dq[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) we[v][k] = ld4(We + (c0 + k) * D_);
    }
    const float t = gsum<LANES>(tpart, mask);
    const float m = rmax[i], inv_den = 1.0f / rden[i];
    const int32_t beg = rowptr[i], end = rowptr[i + 1];
    for (int32_t p = beg; p < end; ++p) {
      const int64_t j = src[p];
      const int32_t e_id = eid[p];
      const float4 ea = ld4(edge_attr + static_cast<int64_t>(e_id) * D_);
      float4 ke[VEC];
      float part = 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int c0 = (v * LANES + l) * 4;
        const float4 e = edge_proj(we[v], ea);
        ke[v] = add4(ld4(qkvs + j * 4 * H + H + c0), e);
        const float4 ve = add4(ld4(qkvs + j * 4 * H + 2 * H + c0), e);
        part += dot4(g[v], ve);
      }
      const float dalpha = gsum<LANES>(part, mask);
      const float alpha = expf(logit[p] - m) * inv_den;
      const float da = alpha * (dalpha - t) * inv_sqrt;
      if (l == 0) {
        wa[e_id] = alpha;
        wda[e_id] = da;
      }
      const float eav[4] = {ea.x, ea.y, ea.z, ea.w};
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        dq[v] = fma4(da, ke[v], dq[v]);
        const float4 de = fma4(alpha, g[v], scale4(da, q[v]));
        const float dev[4] = {de.x, de.y, de.z, de.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // dWe[c0+k][d] += de[k] * a[d]
          dwe[v][k].x = fmaf(dev[k], eav[0], dwe[v][k].x);
          dwe[v][k].y = fmaf(dev[k], eav[1], dwe[v][k].y);
          dwe[v][k].z = fmaf(dev[k], eav[2], dwe[v][k].z);
          dwe[v][k].w = fmaf(dev[k], eav[3], dwe[v][k].w);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) st4(dqkvs + i * 4 * H + (v * LANES + l) * 4, dq[v]);
  }
  // block partial of dWe: fixed-order sum over the block's rows
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      st4(s_red + static_cast<size_t>(rl) * H * D_ + ((v * LANES + l) * 4 + k) * D_, dwe[v][k]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < H * D_; idx += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < ROWS; ++r) a += s_red[static_cast<size_t>(r) * H * D_ + idx];
    dwe_part[static_cast<int64_t>(blockIdx.x) * H * D_ + idx] = a;
  }
}

// TransformerConv backward, source side (transposed CSR):
//   dk_j = sum_{j->i} wda[e] q_i ;  dv_j = sum_{j->i} wa[e] g_i
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
tconv_bwd_src_kernel(const float* __restrict__ qkvs, const int32_t* __restrict__ t_rowptr,
                     const int32_t* __restrict__ t_dst, const int32_t* __restrict__ t_eid,
                     const float* __restrict__ wa, const float* __restrict__ wda, int64_t N,
                     float* __restrict__ dqkvs) {
  constexpr int H = LANES * VEC * 4;
  const int64_t j = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / LANES;
  const int l = threadIdx.x % LANES;
  if (j >= N) return;
  float4 dk[VEC], dv[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dk[v] = dv[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int32_t beg = t_rowptr[j], end = t_rowptr[j + 1];
  for (int32_t p = beg; p < end; ++p) {
    const int64_t i = t_dst[p];
    const int32_t e = t_eid[p];
    const float a = wa[e], da = wda[e];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c0 = (v * LANES + l) * 4;
      dk[v] = fma4(da, ld4(qkvs + i * 4 * H + c0), dk[v]);
      dv[v] = fma4(a, ld4(dqkvs + i * 4 * H + 3 * H + c0), dv[v]);   // g_i lives in the skip slab
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = (v * LANES + l) * 4;
    st4(dqkvs + j * 4 * H + H + c0, dk[v]);
    st4(dqkvs + j * 4 * H + 2 * H + c0, dv[v]);
  }
}

// sum of per-block partials in a fixed order: one block per 8 outputs, 32 lanes stride over the
// partial blocks, then a shuffle tree (the order depends on nblocks only)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ part, int64_t nblocks, int64_t n, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (idx >= n) return;
  float a = 0.f;
  for (int64_t b = lane; b < nblocks; b += 32) a += part[b * n + idx];
  a = warp_sum(a);
  if (lane == 0) out[idx] = a;
}

// ===========================================================================
// NNConv(mean) forward, factorised: out_i = mean_j sum_k hh_e[k] yr[j,k,:] + root_i + bias
// ===========================================================================
__device__ __forceinline__ void edge_hidden(const EdgeMlp1& w1, const float4& ea, float (&hh)[K_]) {
#pragma unroll
  for (int k = 0; k < K_; ++k) {
    float v = w1.b[k];
    v = fmaf(w1.w[k][0], ea.x, v);
    v = fmaf(w1.w[k][1], ea.y, v);
    v = fmaf(w1.w[k][2], ea.z, v);
    v = fmaf(w1.w[k][3], ea.w, v);
    hh[k] = v;            // pre-activation; callers apply relu
  }
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
nnconv_fwd_kernel(const float* __restrict__ yr, const int32_t* __restrict__ rowptr,
                  const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                  const float* __restrict__ edge_attr, const float* __restrict__ W1,
                  const float* __restrict__ b1, const float* __restrict__ bias,
                  int64_t N, float slope, float* __restrict__ out) {
  constexpr int H = LANES * VEC * 4;
  constexpr int LD = (K_ + 2) * H;
  const EdgeMlp1 w1 = load_edge_mlp1(W1, b1);
  const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / LANES;
  const int l = threadIdx.x % LANES;
  if (i >= N) return;
  float4 acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  for (int32_t p = beg; p < end; ++p) {
    const float* yj = yr + static_cast<int64_t>(src[p]) * LD;
    const float4 ea = ld4(edge_attr + static_cast<int64_t>(eid[p]) * D_);
    float hh[K_];
    edge_hidden(w1, ea, hh);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c0 = (v * LANES + l) * 4;
      float4 msg = ld4(yj + K_ * H + c0);
#pragma unroll
      for (int k = 0; k < K_; ++k) msg = fma4(fmaxf(hh[k], 0.f), ld4(yj + k * H + c0), msg);
      acc[v] = add4(acc[v], msg);
    }
  }
  const float inv = 1.0f / static_cast<float>(max(end - beg, 1));
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = (v * LANES + l) * 4;
    const float4 o = add4(fma4(inv, acc[v], ld4(yr + i * LD + (K_ + 1) * H + c0)), ld4(bias + c0));
    st4(out + i * H + c0, lrelu4(o, slope));
  }
}

// NNConv backward, destination side: root slab of dyr = g; per-edge hidden gradient
// dhid[k] = relu'(.) (g/deg . yr[j,k,:]) accumulated into dW1/db1 block partials.
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
nnconv_bwd_dst_kernel(const float* __restrict__ yr, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                      const float* __restrict__ edge_attr, const float* __restrict__ W1,
                      const float* __restrict__ b1, const float* __restrict__ out,
                      const float* __restrict__ dout, int64_t N, float slope,
                      float* __restrict__ dyr, float* __restrict__ w1_part) {
  constexpr int H = LANES * VEC * 4;
  constexpr int LD = (K_ + 2) * H;
  constexpr int ROWS = 256 / LANES;
  constexpr int NW = K_ * D_ + K_;                 // 40 accumulators: dW1 [8][4], db1 [8]
  __shared__ float s_red[ROWS][NW + 1];
  const EdgeMlp1 w1 = load_edge_mlp1(W1, b1);
  const int rl = threadIdx.x / LANES, l = threadIdx.x % LANES;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * ROWS + rl;
  const unsigned mask = group_mask<LANES>();
  float accw[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) accw[k] = 0.f;
  if (i < N) {
    const int32_t beg = rowptr[i], end = rowptr[i + 1];
    const float inv = 1.0f / static_cast<float>(max(end - beg, 1));
    float4 gs[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c0 = (v * LANES + l) * 4;
      const float4 g = lrelu_grad4(ld4(dout + i * H + c0), ld4(out + i * H + c0), slope);
      st4(dyr + i * LD + (K_ + 1) * H + c0, g);
      gs[v] = scale4(inv, g);
    }
    for (int32_t p = beg; p < end; ++p) {
      const float* yj = yr + static_cast<int64_t>(src[p]) * LD;
      const float4 ea = ld4(edge_attr + static_cast<int64_t>(eid[p]) * D_);
      float hh[K_];
      edge_hidden(w1, ea, hh);
#pragma unroll
      for (int k = 0; k < K_; ++k) {
        float part = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) part += dot4(gs[v], ld4(yj + k * H + (v * LANES + l) * 4));
        const float red = gsum<LANES>(part, mask);
        const float dh = hh[k] > 0.f ? red : 0.f;
        accw[k * D_ + 0] = fmaf(dh, ea.x, accw[k * D_ + 0]);
        accw[k * D_ + 1] = fmaf(dh, ea.y, accw[k * D_ + 1]);
        accw[k * D_ + 2] = fmaf(dh, ea.z, accw[k * D_ + 2]);
        accw[k * D_ + 3] = fmaf(dh, ea.w, accw[k * D_ + 3]);
        accw[K_ * D_ + k] += dh;
      }
    }
  }
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < NW; ++k) s_red[rl][k] = accw[k];
  }
  __syncthreads();
  if (threadIdx.x < NW) {
    float a = 0.f;
    for (int r = 0; r < ROWS; ++r) a += s_red[r][threadIdx.x];
    w1_part[static_cast<int64_t>(blockIdx.x) * NW + threadIdx.x] = a;
  }
}

// NNConv backward, source side: dyr[j,k,:] = sum_{j->i} hh_e[k] g_i/deg_i ; slab K: sum g_i/deg_i
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
nnconv_bwd_src_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ t_rowptr,
                      const int32_t* __restrict__ t_dst, const int32_t* __restrict__ t_eid,
                      const float* __restrict__ edge_attr, const float* __restrict__ W1,
                      const float* __restrict__ b1, int64_t N, float* __restrict__ dyr) {
  constexpr int H = LANES * VEC * 4;
  constexpr int LD = (K_ + 2) * H;
  const EdgeMlp1 w1 = load_edge_mlp1(W1, b1);
  const int64_t j = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / LANES;
  const int l = threadIdx.x % LANES;
  if (j >= N) return;
  float4 acc[VEC][K_ + 1];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int k = 0; k <= K_; ++k) acc[v][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int32_t beg = t_rowptr[j], end = t_rowptr[j + 1];
  for (int32_t p = beg; p < end; ++p) {
    const int64_t i = t_dst[p];
    const float inv = 1.0f / static_cast<float>(max(rowptr[i + 1] - rowptr[i], 1));
    const float4 ea = ld4(edge_attr + static_cast<int64_t>(t_eid[p]) * D_);
    float hh[K_];
    edge_hidden(w1, ea, hh);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float4 gs = scale4(inv, ld4(dyr + i * LD + (K_ + 1) * H + (v * LANES + l) * 4));
#pragma unroll
      for (int k = 0; k < K_; ++k) acc[v][k] = fma4(fmaxf(hh[k], 0.f), gs, acc[v][k]);
      acc[v][K_] = add4(acc[v][K_], gs);
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int k = 0; k <= K_; ++k) st4(dyr + j * LD + k * H + (v * LANES + l) * 4, acc[v][k]);
}

// ===========================================================================
// global_mean_pool + MLP head.
//   stage 1 (pool_partial_kernel): block (g, s) sums rows of chunk s of graph g -- 256 threads =
//           H channels x 256/H row lanes, coalesced rows, row lanes combined in a fixed order;
//           S = ceil(avg graph size / 256) chunks so a 10k-node graph spreads over 40 blocks;
//   stage 2 (pool_mlp_fwd_kernel): one warp per graph adds the S partials in order, divides by
//           the node count, and runs Linear - LeakyReLU - (dropout mask) - Linear.
// ===========================================================================
constexpr int kPoolWarps = 4;
constexpr int kMaxH = 256;
constexpr int kPoolMaxSplit = 2 * kNumSMs;

// chunks per graph: enough blocks to fill the machine when there are few big graphs (a single 10k-node graph
// spreads over 296 blocks of ~34 rows), one chunk per graph when there are many small ones
static int pool_split(int64_t N, int64_t B) {
  const int64_t avg = B > 0 ? cdiv(N, B) : 0;
  const int64_t fill = cdiv(static_cast<int64_t>(kPoolMaxSplit), std::max<int64_t>(B, 1));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(fill, kPoolMaxSplit), cdiv(avg, 32))));
}

__device__ __forceinline__ void pool_chunk(const int64_t* __restrict__ gptr, int64_t g, int s, int S,
                                           int64_t& r0, int64_t& r1) {
  const int64_t n0 = gptr[g], n1 = gptr[g + 1];
  const int64_t per = (n1 - n0 + S - 1) / S;
  r0 = min(n0 + s * per, n1);
  r1 = min(r0 + per, n1);
}

// sum of rows [r0, r1) of a row-major [*, H] matrix by one 256-thread block, H % 4 == 0: thread = (row lane, channel
// quad), 128-bit loads, row lanes combined through shared memory in a fixed order.  Result: s_out[0..H) (valid after
// the trailing __syncthreads).  s_acc: 256 float4 of scratch.
__device__ __forceinline__ void pool_block_rowsum(const float* __restrict__ x, int64_t r0, int64_t r1, int H,
                                                  float4* s_acc, float* s_out) {
  const int C4 = H >> 2, RL = 256 / C4;
  const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int64_t r = r0 + rl; r < r1; r += RL) {
    const float4 v = x4[r * C4 + c4];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  s_acc[threadIdx.x] = a;
  __syncthreads();
  if (rl == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < RL; ++k) {
      const float4 v = s_acc[k * C4 + c4];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(s_out)[c4] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
pool_partial_kernel(const float* __restrict__ x, const int64_t* __restrict__ gptr, int H, int S,
                    float* __restrict__ partial) {
  __shared__ float4 s_acc[256];
  __shared__ __align__(16) float s_out[kMaxH];
  const int64_t g = blockIdx.x;
  const int s = blockIdx.y;
  int64_t r0, r1;
  pool_chunk(gptr, g, s, S, r0, r1);
  pool_block_rowsum(x, r0, r1, H, s_acc, s_out);
  if (threadIdx.x < H) partial[(g * S + s) * H + threadIdx.x] = s_out[threadIdx.x];
}

// Head, one graph per TPG threads (TPG = 32: a warp per graph for H <= 32; TPG = 256: a block per
// graph for wide hidden sizes, so the H x H weight is read with coalesced rows and the work of a
// single big graph is not left to one warp).  pooled = sum of the S partials / n.
template <int TPG>
__global__ void __launch_bounds__(kPoolWarps * 32 < TPG ? TPG : kPoolWarps * 32)
pool_mlp_fwd_kernel(const float* __restrict__ partial, int S, const int64_t* __restrict__ gptr, int64_t B, int H,
                    const float* __restrict__ W1, const float* __restrict__ b1,
                    const float* __restrict__ W2, const float* __restrict__ b2,
                    const float* __restrict__ hmask, float* __restrict__ pooled,
                    float* __restrict__ hid, float* __restrict__ out) {
  constexpr int GPB = (kPoolWarps * 32 < TPG ? TPG : kPoolWarps * 32) / TPG;   // graphs per block
  __shared__ __align__(16) float s_p[GPB][kMaxH];
  __shared__ float s_a[GPB][kMaxH];
  __shared__ float4 s_acc[TPG == 256 ? 256 : 1];
  const int gl = threadIdx.x / TPG, t = threadIdx.x % TPG;
  const int lane = threadIdx.x & 31;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * GPB + gl;
  const bool live = g < B;
  if (TPG == 256) {                                    // one graph per block: all threads add the S partial rows
    const int64_t n0 = gptr[g], n1 = gptr[g + 1];
    const float inv = 1.0f / static_cast<float>(max(n1 - n0, static_cast<int64_t>(1)));
    pool_block_rowsum(partial + g * S * H, 0, S, H, s_acc, s_p[0]);
    for (int c = t; c < H; c += TPG) {
      const float a = s_p[0][c] * inv;
      s_p[0][c] = a;
      if (pooled) pooled[g * H + c] = a;
    }
  } else if (live) {
    const int64_t n0 = gptr[g], n1 = gptr[g + 1];
    const float inv = 1.0f / static_cast<float>(max(n1 - n0, static_cast<int64_t>(1)));
    for (int c = t; c < H; c += TPG) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += partial[(g * S + s) * H + c];
      a *= inv;
      s_p[gl][c] = a;
      if (pooled) pooled[g * H + c] = a;
    }
  }
  if (TPG > 32) __syncthreads(); else __syncwarp();
  if (live) {
    if (TPG == 32) {                                   // lane u owns hidden unit u (H <= 32 here: tiny weights)
      for (int u = t; u < H; u += 32) {
        float hv = b1[u];
        for (int c = 0; c < H; ++c) hv = fmaf(W1[u * H + c], s_p[gl][c], hv);
        if (hid) hid[g * H + u] = hv;
        float act = leaky(hv, 0.01f);
        if (hmask) act *= hmask[g * H + u];
        s_a[gl][u] = act;
      }
    } else {                                           // warp w owns units w, w+8, ...; lanes stride the row with
      const int w = t >> 5;                            // 128-bit loads; eight units per round, all their weight-row
      constexpr int NW = TPG / 32, UPR = 8;            // loads issued before the first is used (H >= 64: H / NW >= 8)
      const float4* W4 = reinterpret_cast<const float4*>(W1);
      const float4* p4 = reinterpret_cast<const float4*>(s_p[gl]);
      const int C4 = H >> 2;
      for (int u0 = w; u0 < H; u0 += UPR * NW) {
        float4 wv[UPR][2];
#pragma unroll
        for (int q = 0; q < UPR; ++q)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c4 = lane + 32 * i;
            wv[q][i] = c4 < C4 ? W4[static_cast<int64_t>(u0 + q * NW) * C4 + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        float4 pv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) pv[i] = lane + 32 * i < C4 ? p4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < UPR; ++q) {
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            part = fmaf(wv[q][i].x, pv[i].x, part); part = fmaf(wv[q][i].y, pv[i].y, part);
            part = fmaf(wv[q][i].z, pv[i].z, part); part = fmaf(wv[q][i].w, pv[i].w, part);
          }
          const float sum = warp_sum(part);
          const int u = u0 + q * NW;
          if (lane == 0) {
            const float hv = sum + b1[u];
            if (hid) hid[g * H + u] = hv;
            float act = leaky(hv, 0.01f);
            if (hmask) act *= hmask[g * H + u];
            s_a[gl][u] = act;
          }
        }
      }
    }
  }
  if (TPG > 32) __syncthreads(); else __syncwarp();
  if (live && t < 32) {                                // first warp of the graph: the 3 outputs
    float o[QOT_OUT] = {0.f, 0.f, 0.f};
    for (int u = lane; u < H; u += 32) {
      const float act = s_a[gl][u];
#pragma unroll
      for (int k = 0; k < QOT_OUT; ++k) o[k] = fmaf(W2[k * H + u], act, o[k]);
    }
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) o[k] = warp_sum(o[k]);
    if (lane < QOT_OUT) out[g * QOT_OUT + lane] = (lane == 0 ? o[0] : lane == 1 ? o[1] : o[2]) + b2[lane];
  }
}

// backward, per graph: dhid, act (for the weight gradients) and dpool [B,H] = W1^T dhid / n_g
template <int TPG>
__global__ void __launch_bounds__(kPoolWarps * 32 < TPG ? TPG : kPoolWarps * 32)
pool_mlp_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ hid,
                    const float* __restrict__ hmask, const int64_t* __restrict__ gptr, int64_t B, int H,
                    const float* __restrict__ W1, const float* __restrict__ W2,
                    float* __restrict__ dpool, float* __restrict__ dhid, float* __restrict__ act) {
  constexpr int GPB = (kPoolWarps * 32 < TPG ? TPG : kPoolWarps * 32) / TPG;
  __shared__ float s_d[GPB][kMaxH];
  const int gl = threadIdx.x / TPG, t = threadIdx.x % TPG;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * GPB + gl;
  const bool live = g < B;
  if (live) {
    const float d0 = dout[g * QOT_OUT + 0], d1 = dout[g * QOT_OUT + 1], d2 = dout[g * QOT_OUT + 2];
    for (int u = t; u < H; u += TPG) {
      const float hv = hid[g * H + u];
      const float mk = hmask ? hmask[g * H + u] : 1.f;
      const float dact = W2[0 * H + u] * d0 + W2[1 * H + u] * d1 + W2[2 * H + u] * d2;
      const float dh = dact * mk * (hv > 0.f ? 1.f : 0.01f);
      s_d[gl][u] = dh;
      dhid[g * H + u] = dh;
      act[g * H + u] = leaky(hv, 0.01f) * mk;
    }
  }
  if (TPG > 32) __syncthreads(); else __syncwarp();
  if (live) {
    const int64_t n0 = gptr[g], n1 = gptr[g + 1];
    const float inv = 1.0f / static_cast<float>(max(n1 - n0, static_cast<int64_t>(1)));
    for (int c = t; c < H; c += TPG) {                 // W1[u*H + c]: coalesced over c
      float dp = 0.f;
      for (int u = 0; u < H; ++u) dp = fmaf(W1[u * H + c], s_d[gl][u], dp);
      dpool[g * H + c] = dp * inv;
    }
  }
}

// dx[n,:] = dpool[graph(n),:]: block (g, s) broadcasts over its chunk of rows
__global__ void __launch_bounds__(256)
pool_bwd_rows_kernel(const float* __restrict__ dpool, const int64_t* __restrict__ gptr, int H, int S,
                     float* __restrict__ dx) {
  const int64_t g = blockIdx.x;
  const int s = blockIdx.y;
  const int c = threadIdx.x % H, rl = threadIdx.x / H, RL = 256 / H;
  int64_t r0, r1;
  pool_chunk(gptr, g, s, S, r0, r1);
  const float v = dpool[g * H + c];
  for (int64_t r = r0 + rl; r < r1; r += RL) dx[r * H + c] = v;
}

// stand-alone global_mean_pool (PyG layer used on its own, topological_training/models.py:61 when
// the reference's models.py is kept): partials -> mean; backward broadcasts dpooled / n_g
__global__ void pool_mean_final_kernel(const float* __restrict__ partial, int S, const int64_t* __restrict__ gptr,
                                       int64_t B, int H, float* __restrict__ pooled) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= B * H) return;
  const int64_t g = i / H;
  const int c = static_cast<int>(i % H);
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(g * S + s) * H + c];
  pooled[i] = a / static_cast<float>(max(gptr[g + 1] - gptr[g], static_cast<int64_t>(1)));
}
// the same with one block per graph (few big graphs: S is large, the serial loop above would be latency-bound)
__global__ void __launch_bounds__(256)
pool_mean_final_block_kernel(const float* __restrict__ partial, int S, const int64_t* __restrict__ gptr, int H,
                             float* __restrict__ pooled) {
  __shared__ float4 s_acc[256];
  __shared__ __align__(16) float s_out[kMaxH];
  const int64_t g = blockIdx.x;
  pool_block_rowsum(partial + g * S * H, 0, S, H, s_acc, s_out);
  if (threadIdx.x < H)
    pooled[g * H + threadIdx.x] = s_out[threadIdx.x] / static_cast<float>(max(gptr[g + 1] - gptr[g], static_cast<int64_t>(1)));
}
__global__ void pool_scale_kernel(const float* __restrict__ dpooled, const int64_t* __restrict__ gptr, int64_t B,
                                  int H, float* __restrict__ dscaled) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= B * H) return;
  const int64_t g = i / H;
  dscaled[i] = dpooled[i] / static_cast<float>(max(gptr[g + 1] - gptr[g], static_cast<int64_t>(1)));
}

// ---------------------------------------------------------------------------
// dispatch on H
// ---------------------------------------------------------------------------
#define QOT_DISPATCH_H(H, ...)                                          \
  switch (H) {                                                          \
    case 16:  { constexpr int LANES = 4,  VEC = 1; __VA_ARGS__; } break;       \
    case 32:  { constexpr int LANES = 8,  VEC = 1; __VA_ARGS__; } break;       \
    case 64:  { constexpr int LANES = 16, VEC = 1; __VA_ARGS__; } break;       \
    case 128: { constexpr int LANES = 32, VEC = 1; __VA_ARGS__; } break;       \
    case 256: { constexpr int LANES = 32, VEC = 2; __VA_ARGS__; } break;       \
    default:                                                            \
      set_error("hidden width %lld unsupported (16, 32, 64, 128, 256)", (long long)(H)); \
      return QOT_E_BADARG;                                              \
  }

}  // namespace qot

using namespace qot;

static bool h_ok(int64_t H) { return H == 16 || H == 32 || H == 64 || H == 128 || H == 256; }

extern "C" int qot_tconv_fwd(const float* qkvs, const int32_t* rowptr, const int32_t* src,
                             const int32_t* eid, const float* edge_attr, const float* We,
                             int64_t N, int64_t H, float slope, float* out, float* logit,
                             float* rmax, float* rden, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && h_ok(H), "qot_tconv_fwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  if (N == 0) return QOT_OK;
  QOT_REQUIRE(qkvs && rowptr && src && eid && edge_attr && We && out, "qot_tconv_fwd: null argument");
  QOT_REQUIRE((rmax == nullptr) == (rden == nullptr), "qot_tconv_fwd: rmax and rden go together");
  QOT_DISPATCH_H(H, (tconv_fwd_kernel<LANES, VEC><<<static_cast<unsigned>(cdiv(N * LANES, 256)), 256, 0, stream>>>(
                        qkvs, rowptr, src, eid, edge_attr, We, N, slope, out, logit, rmax, rden)));
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

static int64_t rows_per_block(int64_t H) { return 256 / std::min<int64_t>(H / 4, 32); }

extern "C" size_t qot_tconv_bwd_workspace_bytes(int64_t N, int64_t E, int64_t H) {
  if (N < 0 || E < 0 || !h_ok(H)) return 0;
  const int64_t nb = cdiv(std::max<int64_t>(N, 1), rows_per_block(H));
  return 2 * align_up(static_cast<size_t>(std::max<int64_t>(E, 1)) * 4) +
         align_up(static_cast<size_t>(nb) * H * D_ * 4) + 256;
}

extern "C" int qot_tconv_bwd(const float* qkvs, const int32_t* rowptr, const int32_t* src,
                             const int32_t* eid, const int32_t* t_rowptr, const int32_t* t_dst,
                             const int32_t* t_eid, const float* edge_attr, const float* We,
                             const float* out, const float* dout, const float* logit,
                             const float* rmax, const float* rden, int64_t N, int64_t E, int64_t H,
                             float slope, float* dqkvs, float* dWe,
                             void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && E >= 0 && h_ok(H), "qot_tconv_bwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  QOT_REQUIRE(slope != 0.f, "qot_tconv_bwd: slope 0 is not invertible");
  QOT_REQUIRE(dWe && We, "qot_tconv_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_tconv_bwd_workspace_bytes(N, E, H), "qot_tconv_bwd: workspace too small");
  if (N == 0) {
    QOT_CUDA(cudaMemsetAsync(dWe, 0, H * D_ * 4, stream));
    return QOT_OK;
  }
  QOT_REQUIRE(qkvs && rowptr && src && eid && t_rowptr && t_dst && t_eid && edge_attr && out && dout &&
                  logit && rmax && rden && dqkvs, "qot_tconv_bwd: null argument");
  Carver c(ws);
  float* wa = c.take<float>(std::max<int64_t>(E, 1));
  float* wda = c.take<float>(std::max<int64_t>(E, 1));
  const int64_t rpb = rows_per_block(H), nb = cdiv(N, rpb);
  float* part = c.take<float>(nb * H * D_);
  const size_t smem = static_cast<size_t>(rpb) * H * D_ * sizeof(float);
  QOT_DISPATCH_H(H, {
    if (smem > 48 * 1024)
      QOT_CUDA(cudaFuncSetAttribute(tconv_bwd_dst_kernel<LANES, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tconv_bwd_dst_kernel<LANES, VEC><<<static_cast<unsigned>(nb), 256, smem, stream>>>(
        qkvs, rowptr, src, eid, edge_attr, We, out, dout, logit, rmax, rden, N, slope, dqkvs, wa, wda, part);
  });
  QOT_LAUNCH_CHECK();
  QOT_DISPATCH_H(H, (tconv_bwd_src_kernel<LANES, VEC><<<static_cast<unsigned>(cdiv(N * LANES, 256)), 256, 0, stream>>>(
                        qkvs, t_rowptr, t_dst, t_eid, wa, wda, N, dqkvs)));
  QOT_LAUNCH_CHECK();
  sum_partials_kernel<<<static_cast<unsigned>(cdiv(H * D_, 8)), 256, 0, stream>>>(part, nb, H * D_, dWe);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_nnconv_fwd(const float* yr, const int32_t* rowptr, const int32_t* src,
                              const int32_t* eid, const float* edge_attr, const float* W1,
                              const float* b1, const float* bias, int64_t N, int64_t H, float slope,
                              float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && h_ok(H), "qot_nnconv_fwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  if (N == 0) return QOT_OK;
  QOT_REQUIRE(yr && rowptr && src && eid && edge_attr && W1 && b1 && bias && out, "qot_nnconv_fwd: null argument");
  QOT_DISPATCH_H(H, (nnconv_fwd_kernel<LANES, VEC><<<static_cast<unsigned>(cdiv(N * LANES, 256)), 256, 0, stream>>>(
                        yr, rowptr, src, eid, edge_attr, W1, b1, bias, N, slope, out)));
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_nnconv_bwd_workspace_bytes(int64_t N, int64_t E, int64_t H) {
  if (N < 0 || E < 0 || !h_ok(H)) return 0;
  const int64_t nb = cdiv(std::max<int64_t>(N, 1), rows_per_block(H));
  return align_up(static_cast<size_t>(nb) * 40 * 4) + align_up(40 * 4) +
         qot_colsum_workspace_bytes(std::max<int64_t>(N, 1), H) + 256;
}

extern "C" int qot_nnconv_bwd(const float* yr, const int32_t* rowptr, const int32_t* src,
                              const int32_t* eid, const int32_t* t_rowptr, const int32_t* t_dst,
                              const int32_t* t_eid, const float* edge_attr, const float* W1,
                              const float* b1, const float* out, const float* dout,
                              int64_t N, int64_t E, int64_t H, float slope,
                              float* dyr, float* dW1, float* db1, float* dbias,
                              void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && E >= 0 && h_ok(H), "qot_nnconv_bwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  QOT_REQUIRE(dW1 && db1 && dbias && W1 && b1, "qot_nnconv_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_nnconv_bwd_workspace_bytes(N, E, H), "qot_nnconv_bwd: workspace too small");
  if (N == 0) {
    QOT_CUDA(cudaMemsetAsync(dW1, 0, K_ * D_ * 4, stream));
    QOT_CUDA(cudaMemsetAsync(db1, 0, K_ * 4, stream));
    QOT_CUDA(cudaMemsetAsync(dbias, 0, H * 4, stream));
    return QOT_OK;
  }
  QOT_REQUIRE(yr && rowptr && src && eid && t_rowptr && t_dst && t_eid && edge_attr && out && dout && dyr,
              "qot_nnconv_bwd: null argument");
  Carver c(ws);
  const int64_t rpb = rows_per_block(H), nb = cdiv(N, rpb);
  float* part = c.take<float>(nb * 40);
  float* w1sum = c.take<float>(40);
  const size_t csb = qot_colsum_workspace_bytes(N, H);
  void* csws = c.take<char>(csb);
  QOT_DISPATCH_H(H, (nnconv_bwd_dst_kernel<LANES, VEC><<<static_cast<unsigned>(nb), 256, 0, stream>>>(
                        yr, rowptr, src, eid, edge_attr, W1, b1, out, dout, N, slope, dyr, part)));
  QOT_LAUNCH_CHECK();
  QOT_DISPATCH_H(H, (nnconv_bwd_src_kernel<LANES, VEC><<<static_cast<unsigned>(cdiv(N * LANES, 256)), 256, 0, stream>>>(
                        rowptr, t_rowptr, t_dst, t_eid, edge_attr, W1, b1, N, dyr)));
  QOT_LAUNCH_CHECK();
  sum_partials_kernel<<<5, 256, 0, stream>>>(part, nb, 40, w1sum);
  QOT_LAUNCH_CHECK();
  QOT_CUDA(cudaMemcpyAsync(dW1, w1sum, K_ * D_ * 4, cudaMemcpyDeviceToDevice, stream));
  QOT_CUDA(cudaMemcpyAsync(db1, w1sum + K_ * D_, K_ * 4, cudaMemcpyDeviceToDevice, stream));
  // bias gradient = column sums of g, which lives in the root slab of dyr
  return qot_colsum(dyr + (K_ + 1) * H, (K_ + 2) * H, N, H, dbias, csws, csb, stream_);
}

extern "C" size_t qot_pool_mlp_fwd_workspace_bytes(int64_t N, int64_t B, int64_t H) {
  if (N < 0 || B < 0 || H <= 0) return 0;
  return align_up(static_cast<size_t>(std::max<int64_t>(B, 1)) * pool_split(N, B) * H * 4) + 256;
}

extern "C" int qot_pool_mlp_fwd(const float* x, const int64_t* gptr, int64_t N, int64_t B, int64_t H,
                                const float* W1, const float* b1, const float* W2, const float* b2,
                                const float* hmask, float* pooled, float* hid, float* out,
                                void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(B >= 0 && N >= 0 && h_ok(H), "qot_pool_mlp_fwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  if (B == 0) return QOT_OK;
  QOT_REQUIRE(gptr && W1 && b1 && W2 && b2 && out && (N == 0 || x), "qot_pool_mlp_fwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_pool_mlp_fwd_workspace_bytes(N, B, H), "qot_pool_mlp_fwd: workspace too small");
  QOT_REQUIRE(B <= 0x7fffffffll, "qot_pool_mlp_fwd: too many graphs for one launch");
  QOT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W1)) & 15) == 0,
              "qot_pool_mlp_fwd: x and W1 must be 16-byte aligned (128-bit loads)");
  const int S = pool_split(N, B);
  float* partial = static_cast<float*>(ws);
  pool_partial_kernel<<<dim3(static_cast<unsigned>(B), S), 256, 0, stream>>>(x, gptr, static_cast<int>(H), S, partial);
  QOT_LAUNCH_CHECK();
  if (H <= 32)
    pool_mlp_fwd_kernel<32><<<static_cast<unsigned>(cdiv(B, kPoolWarps)), kPoolWarps * 32, 0, stream>>>(
        partial, S, gptr, B, static_cast<int>(H), W1, b1, W2, b2, hmask, pooled, hid, out);
  else
    pool_mlp_fwd_kernel<256><<<static_cast<unsigned>(B), 256, 0, stream>>>(
        partial, S, gptr, B, static_cast<int>(H), W1, b1, W2, b2, hmask, pooled, hid, out);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_pool_mlp_bwd_workspace_bytes(int64_t B, int64_t H) {
  if (B < 0 || H <= 0) return 0;
  const int64_t b = std::max<int64_t>(B, 1);
  return 3 * align_up(static_cast<size_t>(b) * H * 4) + qot_wgrad_workspace_bytes(b, H, H) +
         qot_wgrad_workspace_bytes(b, QOT_OUT, H) + qot_colsum_workspace_bytes(b, H) + 256;
}

extern "C" int qot_pool_mlp_bwd(const float* dout, const float* pooled, const float* hid,
                                const float* hmask, const int64_t* gptr, int64_t N, int64_t B,
                                int64_t H, const float* W1, const float* W2, float* dx,
                                float* dW1, float* db1, float* dW2, float* db2,
                                void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(B > 0 && N >= 0 && h_ok(H), "qot_pool_mlp_bwd: bad shape");
  QOT_REQUIRE(dout && pooled && hid && gptr && W1 && W2 && dW1 && db1 && dW2 && db2 && (N == 0 || dx),
              "qot_pool_mlp_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_pool_mlp_bwd_workspace_bytes(B, H), "qot_pool_mlp_bwd: workspace too small");
  Carver c(ws);
  float* dhid = c.take<float>(B * H);
  float* act = c.take<float>(B * H);
  float* dpool = c.take<float>(B * H);
  const size_t w1b = qot_wgrad_workspace_bytes(B, H, H), w2b = qot_wgrad_workspace_bytes(B, QOT_OUT, H);
  const size_t csb = qot_colsum_workspace_bytes(B, H);
  void* w1ws = c.take<char>(w1b);
  void* w2ws = c.take<char>(w2b);
  void* csws = c.take<char>(csb);
  if (H <= 32)
    pool_mlp_bwd_kernel<32><<<static_cast<unsigned>(cdiv(B, kPoolWarps)), kPoolWarps * 32, 0, stream>>>(
        dout, hid, hmask, gptr, B, static_cast<int>(H), W1, W2, dpool, dhid, act);
  else
    pool_mlp_bwd_kernel<256><<<static_cast<unsigned>(B), 256, 0, stream>>>(
        dout, hid, hmask, gptr, B, static_cast<int>(H), W1, W2, dpool, dhid, act);
  QOT_LAUNCH_CHECK();
  if (N > 0) {
    const int S = pool_split(N, B);
    pool_bwd_rows_kernel<<<dim3(static_cast<unsigned>(B), S), 256, 0, stream>>>(dpool, gptr, static_cast<int>(H), S, dx);
    QOT_LAUNCH_CHECK();
  }
  int rc;
  if ((rc = qot_wgrad(dhid, H, pooled, H, B, H, H, dW1, H, w1ws, w1b, stream_))) return rc;
  if ((rc = qot_colsum(dhid, H, B, H, db1, csws, csb, stream_))) return rc;
  if ((rc = qot_wgrad(dout, QOT_OUT, act, H, B, QOT_OUT, H, dW2, H, w2ws, w2b, stream_))) return rc;
  if ((rc = qot_colsum(dout, QOT_OUT, B, QOT_OUT, db2, csws, csb, stream_))) return rc;
  return QOT_OK;
}

extern "C" size_t qot_mean_pool_workspace_bytes(int64_t N, int64_t B, int64_t H) {
  return qot_pool_mlp_fwd_workspace_bytes(N, B, H) + align_up(static_cast<size_t>(std::max<int64_t>(B, 1)) * H * 4);
}

extern "C" int qot_mean_pool_fwd(const float* x, const int64_t* gptr, int64_t N, int64_t B, int64_t H,
                                 float* pooled, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(B >= 0 && N >= 0 && h_ok(H), "qot_mean_pool_fwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  if (B == 0) return QOT_OK;
  QOT_REQUIRE(gptr && pooled && (N == 0 || x), "qot_mean_pool_fwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_mean_pool_workspace_bytes(N, B, H), "qot_mean_pool_fwd: workspace too small");
  QOT_REQUIRE(B <= 0x7fffffffll, "qot_mean_pool_fwd: too many graphs for one launch");
  QOT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "qot_mean_pool_fwd: x must be 16-byte aligned (128-bit loads)");
  const int S = pool_split(N, B);
  float* partial = static_cast<float*>(ws);
  pool_partial_kernel<<<dim3(static_cast<unsigned>(B), S), 256, 0, stream>>>(x, gptr, static_cast<int>(H), S, partial);
  QOT_LAUNCH_CHECK();
  if (S > 8)
    pool_mean_final_block_kernel<<<static_cast<unsigned>(B), 256, 0, stream>>>(partial, S, gptr, static_cast<int>(H), pooled);
  else
    pool_mean_final_kernel<<<static_cast<unsigned>(cdiv(B * H, 256)), 256, 0, stream>>>(partial, S, gptr, B, static_cast<int>(H), pooled);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_mean_pool_bwd(const float* dpooled, const int64_t* gptr, int64_t N, int64_t B, int64_t H,
                                 float* dx, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(B >= 0 && N >= 0 && h_ok(H), "qot_mean_pool_bwd: H must be 16/32/64/128/256 (got %lld)", (long long)H);
  if (B == 0 || N == 0) return QOT_OK;
  QOT_REQUIRE(dpooled && gptr && dx, "qot_mean_pool_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_mean_pool_workspace_bytes(N, B, H), "qot_mean_pool_bwd: workspace too small");
  float* dscaled = static_cast<float*>(ws);
  pool_scale_kernel<<<static_cast<unsigned>(cdiv(B * H, 256)), 256, 0, stream>>>(dpooled, gptr, B, static_cast<int>(H), dscaled);
  QOT_LAUNCH_CHECK();
  const int S = pool_split(N, B);
  pool_bwd_rows_kernel<<<dim3(static_cast<unsigned>(B), S), 256, 0, stream>>>(dscaled, gptr, static_cast<int>(H), S, dx);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
