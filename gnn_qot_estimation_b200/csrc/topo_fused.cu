// TopologicalGNN, one thread block per graph: the whole forward (Embedding -> TransformerConv -> LeakyReLU
// -> NNConv(mean) -> LeakyReLU -> global_mean_pool -> MLP; topological_training/models.py:43-64) in one
// launch, and the whole backward (forward recomputed from the inputs, then every parameter gradient) in
// another.  Nothing couples the graphs of a batch in this model, so all activations of a graph live in
// shared memory (n <= 96 nodes: ~1.3 KB per node, 60 B per edge); the 4 092 parameters sit beside them
// (with transposed copies so that every inner loop reads consecutive banks) and, in the backward, so does
// a gradient accumulator of the same size plus the embedding-table gradient.  Blocks are persistent over
// graphs; every gradient element is owned by one thread and summed over nodes / edges / graphs in a fixed
// order; per-block partials go to global memory and are summed in block order by a second kernel:
// bit-reproducible for a given grid.  NNConv in factorised form (T_j = h1_j [P_0..P_7 | P_b]),
// the per-destination softmax in edge order.  The math, line by line, is oracle/topo_fused_math.py.
// Reference shape only: hidden 16, edge_dim 4, edge MLP 4 -> 8 -> 256, out 3; dropout must be inactive.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int TF_H = 16, TF_K = 8, TF_D = 4, TF_KP = 9, TF_T = TF_KP * TF_H;   // T row: 144 floats
// flat parameter / gradient layout (floats); W2 | b2 are held as P[c][k*16+o] in shared memory
constexpr int oWq = 0, obq = 256, oWk = 272, obk = 528, oWv = 544, obv = 800, oWs = 816, obs = 1072, oWe = 1088,
              oW1 = 1152, ob1 = 1184, oW2 = 1192, ob2 = 3240, oWroot = 3496, obias2 = 3752, oWm1 = 3768,
              obm1 = 4024, oWm2 = 4040, obm2 = 4088, kTfParams = 4092;
constexpr int kTfThreads = 128;
#ifdef QOT_LP_TRACE
__device__ long long* g_tf_trace = nullptr;               // debug build: 24 clock stamps per block (scripts/trace_topo_fused.py)
#define TF_STAMP(k) do { if (threadIdx.x == 0 && g_tf_trace) g_tf_trace[blockIdx.x * 24 + (k)] = clock64(); } while (0)
#else
#define TF_STAMP(k) do {} while (0)
#endif
constexpr float kTfSlope = 0.01f;

constexpr int kTfPrepared = kTfParams + 6 * 256;          // flat (P layout at oW2) | WqT WkT WvT WsT WrootT Wm1T
struct TfPtrs {                  // par..Wm1T: the prepared GLOBAL buffer (16 KB, L1-resident); the rest is shared memory
  const float *par, *WqT, *WkT, *WvT, *WsT, *WrootT, *Wm1T;    // par: flat layout, with P at oW2 (oW2..ob2+256)
  float *grad, *gemb;                                      // backward only
  float *X, *Q, *K, *V, *H1, *B1, *B2, *G1, *G2, *G3, *T;  // node arrays (G*: backward only)
  float *attr, *hid, *alpha, *dlogit, *dhid;               // edge arrays (dlogit, dhid: backward only)
  int *rowd, *rows;                                        // CSR by destination / by source: [nmax + 1]
  unsigned short *src, *dst, *permd, *perms;               // [emax]
  float *small;                                            // pool[16] pre1[16] z1[16] out[4] dpool[16] dpre1[16] dout[4]
};

// Dropout of one graph (topological_training/models.py:55,59 on the node features, :36-41 inside the head): the
// masks are kernel INPUTS -- one byte per element, drawn by the caller with torch's generator (so CUDA-graph replays
// advance it and a test can replay the same masks through the oracle) -- 1 keeps the element and scales it by
// 1 / (1 - p).  m1 / m2: [n,16] after conv1 / conv2, m3: [16] inside the head; all NULL = no dropout.
struct TfDrop {
  const uint8_t *m1, *m2, *m3;
  float scale, scale_h;
};
__device__ __forceinline__ TfDrop tf_drop_of(const uint8_t* mask, float scale, float scale_h, int64_t Ntot, int64_t n0, int64_t g) {
  TfDrop d;
  d.m1 = mask ? mask + n0 * TF_H : nullptr;
  d.m2 = mask ? mask + (Ntot + n0) * TF_H : nullptr;
  d.m3 = mask ? mask + 2 * Ntot * TF_H + g * TF_H : nullptr;
  d.scale = scale; d.scale_h = scale_h;
  return d;
}
__device__ __forceinline__ float tf_keep(const uint8_t* m, int idx, float scale) { return m ? (m[idx] ? scale : 0.f) : 1.f; }

__host__ __device__ inline size_t tf_smem_bytes(int nmax, int emax, int num_nodes, bool bwd) {
  size_t f = 0;
  if (bwd) f += kTfParams + static_cast<size_t>(num_nodes) * TF_H;
  f += static_cast<size_t>(nmax) * (TF_H * (bwd ? 10 : 7) + TF_T);
  f += static_cast<size_t>(emax) * (TF_D + TF_K + (bwd ? 2 + TF_K : 1));
  f += 96;
  size_t b = f * 4 + 2 * static_cast<size_t>(nmax + 1) * 4 + 4 * static_cast<size_t>(emax) * 2;
  return (b + 15) & ~static_cast<size_t>(15);
}

__device__ inline TfPtrs tf_carve(char* base, const float* __restrict__ prep, int nmax, int emax, int num_nodes, bool bwd) {
  TfPtrs p;
  float* f = reinterpret_cast<float*>(base);
  auto take = [&](size_t n) { float* r = f; f += n; return r; };
  p.par = prep;
  p.WqT = prep + kTfParams; p.WkT = p.WqT + 256; p.WvT = p.WkT + 256; p.WsT = p.WvT + 256; p.WrootT = p.WsT + 256;
  p.Wm1T = p.WrootT + 256;
  p.grad = bwd ? take(kTfParams) : nullptr;
  p.gemb = bwd ? take(static_cast<size_t>(num_nodes) * TF_H) : nullptr;
  p.X = take(nmax * TF_H); p.Q = take(nmax * TF_H); p.K = take(nmax * TF_H); p.V = take(nmax * TF_H);
  p.H1 = take(nmax * TF_H); p.B1 = take(nmax * TF_H); p.B2 = take(nmax * TF_H);
  p.G1 = bwd ? take(nmax * TF_H) : nullptr; p.G2 = bwd ? take(nmax * TF_H) : nullptr; p.G3 = bwd ? take(nmax * TF_H) : nullptr;
  p.T = take(static_cast<size_t>(nmax) * TF_T);
  p.attr = take(static_cast<size_t>(emax) * TF_D); p.hid = take(static_cast<size_t>(emax) * TF_K); p.alpha = take(emax);
  p.dlogit = bwd ? take(emax) : nullptr;
  p.dhid = bwd ? take(static_cast<size_t>(emax) * TF_K) : nullptr;
  p.small = take(96);
  int* ip = reinterpret_cast<int*>(f);
  p.rowd = ip; ip += nmax + 1;
  p.rows = ip; ip += nmax + 1;
  unsigned short* sp = reinterpret_cast<unsigned short*>(ip);
  p.src = sp; sp += emax; p.dst = sp; sp += emax; p.permd = sp; sp += emax; p.perms = sp;
  return p;
}

__device__ __forceinline__ float tf_lk(float v) { return v > 0.f ? v : kTfSlope * v; }
__device__ __forceinline__ float tf_dlk_from_act(float act) { return act > 0.f ? 1.f : kTfSlope; }   // act = leaky(pre): same sign
__device__ __forceinline__ float tf_sum16(float v) {                       // over an aligned 16-lane group
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o, 16);
  return v;
}

// parameters -> the prepared buffer (flat copy, P rearrangement of W2 | b2, transposed 16x16 matrices)
__global__ void __launch_bounds__(256)
topo_fused_prepare_kernel(const float* __restrict__ flat, float* __restrict__ prep) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = t0; i < kTfParams; i += nt)
    if (i < oW2 || i >= oWroot) prep[i] = flat[i];
  // P[c][k*16+o] = W2[(c*16+o)*8 + k] (k < 8), P[c][8*16+o] = b2[c*16+o]; stored at oW2, row stride 144
  for (int i = t0; i < TF_H * TF_T; i += nt) {
    const int c = i / TF_T, r = i % TF_T, k = r / TF_H, o = r % TF_H;
    prep[oW2 + i] = k < TF_K ? flat[oW2 + (c * TF_H + o) * TF_K + k] : flat[ob2 + c * TF_H + o];
  }
  float* t = prep + kTfParams;
  for (int i = t0; i < 256; i += nt) {
    const int a = i >> 4, b = i & 15;                  // T[b][a] = W[a][b]
    t[b * 16 + a] = flat[oWq + i]; t[256 + b * 16 + a] = flat[oWk + i]; t[512 + b * 16 + a] = flat[oWv + i];
    t[768 + b * 16 + a] = flat[oWs + i]; t[1024 + b * 16 + a] = flat[oWroot + i]; t[1280 + b * 16 + a] = flat[oWm1 + i];
  }
}

// One graph, forward.  Leaves in shared memory everything the backward needs: X Q K V H1 (leaky(O1)),
// B1 = H2 (leaky(O2)), T, attr, hid, alpha, both CSRs, small[] = pool | pre1 | z1 | out.
// Saved forward state of one graph (training): [Q | K | V | H1 | H2 | T] per node range, [hid | alpha] per edge
// range, pool | pre1 | z1 | out per graph.  kTfSaveNode floats per node, kTfSaveEdge per edge, kTfSaveGraph per graph.
constexpr int kTfSaveNode = 5 * TF_H + TF_T, kTfSaveEdge = TF_K + 1, kTfSaveGraph = 64;

// `save` (nullable): the forward state is also written there; `restore` (nullable): the state is READ from there
// instead of being computed (the backward kernel, when the forward kernel of the same step saved it)
__device__ void tf_graph_forward(const TfPtrs& p, const float* __restrict__ emb, const int64_t* __restrict__ node_ids,
                                 const int64_t* __restrict__ esrc, const int64_t* __restrict__ edst,
                                 const float* __restrict__ eattr, int64_t n0, int n, int64_t e0, int E, int num_nodes,
                                 float* __restrict__ save_n, float* __restrict__ save_e, float* __restrict__ save_g,
                                 const float* __restrict__ rest_n, const float* __restrict__ rest_e,
                                 const float* __restrict__ rest_g, int32_t* __restrict__ status, const TfDrop& dr) {
  const int tid = threadIdx.x;
  const float* __restrict__ par = p.par;
  // ---- inputs.  Out-of-range ids / endpoints are clamped for memory safety AND reported (status bit 2: the
  // reference's embedding lookup / scatter would raise on them); ops.check_status() turns the bit into an error
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int i = idx >> 4, c = idx & 15;
    long long id = node_ids[n0 + i];
    if (id < 0 || id >= num_nodes) {
      if (c == 0) atomicOr(status, 2);
      id = id < 0 ? 0 : num_nodes - 1;
    }
    p.X[idx] = emb[id * TF_H + c];
  }
  for (int e = tid; e < E; e += kTfThreads) {            // endpoints outside the graph's node range: an edge_ptr that
    const long long a = esrc[e0 + e] - n0, b = edst[e0 + e] - n0;   // does not describe edge_index
    const bool ok = a >= 0 && a < n && b >= 0 && b < n;
    if (!ok) atomicOr(status, 2);
    p.src[e] = static_cast<unsigned short>(a >= 0 && a < n ? a : 0);
    p.dst[e] = static_cast<unsigned short>(b >= 0 && b < n ? b : 0);
  }
  for (int idx = tid; idx < E * TF_D; idx += kTfThreads) p.attr[idx] = eattr[e0 * TF_D + idx];
  __syncthreads();
  TF_STAMP(1);
  // ---- both CSRs, stable (edge order inside a row): one edge per thread -- its rank among the earlier
  // edges of the same row; row sizes by integer atomics (order-free); offsets by a warp scan
  for (int i = tid; i <= n; i += kTfThreads) { p.rowd[i] = 0; p.rows[i] = 0; }
  __syncthreads();
  int rank_d[4], rank_s[4];                             // edges tid, tid+128, ... (emax <= 512 on this path: checked on the host)
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = tid + u * kTfThreads;
    rank_d[u] = rank_s[u] = 0;
    if (e < E) {
      const int d = p.dst[e], sr = p.src[e];
      int rd = 0, rs = 0;
      for (int q = 0; q < e; ++q) { rd += p.dst[q] == d; rs += p.src[q] == sr; }
      rank_d[u] = rd; rank_s[u] = rs;
      atomicAdd(&p.rowd[d + 1], 1);
      atomicAdd(&p.rows[sr + 1], 1);
    }
  }
  __syncthreads();
  if (tid < 64) {                                       // warp 0: rowd, warp 1: rows; inclusive scan in chunks of 32
    int* row = tid < 32 ? p.rowd : p.rows;
    const int ln = tid & 31;
    int carry = 0;
    for (int b0 = 0; b0 <= n; b0 += 32) {
      int v = b0 + ln <= n ? row[b0 + ln] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, v, o);
        if (ln >= o) v += up;
      }
      if (b0 + ln <= n) row[b0 + ln] = v + carry;
      carry += __shfl_sync(kFull, v, 31);
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = tid + u * kTfThreads;
    if (e < E) {
      p.permd[p.rowd[p.dst[e]] + rank_d[u]] = static_cast<unsigned short>(e);
      p.perms[p.rows[p.src[e]] + rank_s[u]] = static_cast<unsigned short>(e);
    }
  }
  TF_STAMP(2);
  if (rest_n) {                                          // restore instead of recomputing
    const int nh = n * TF_H;
    for (int idx = tid; idx < nh; idx += kTfThreads) {
      p.Q[idx] = rest_n[idx]; p.K[idx] = rest_n[nh + idx]; p.V[idx] = rest_n[2 * nh + idx];
      p.H1[idx] = rest_n[3 * nh + idx]; p.B1[idx] = rest_n[4 * nh + idx];
    }
    for (int idx = tid; idx < n * TF_T; idx += kTfThreads) p.T[idx] = rest_n[5 * nh + idx];
    for (int idx = tid; idx < E * TF_K; idx += kTfThreads) p.hid[idx] = rest_e[idx];
    for (int idx = tid; idx < E; idx += kTfThreads) p.alpha[idx] = rest_e[E * TF_K + idx];
    if (tid < 52) p.small[tid] = rest_g[tid];
    __syncthreads();
    return;
  }
  // ---- node projections q k v (skip goes straight into O1, held in B2 for now)
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int i = idx >> 4, c = idx & 15;
    float q = par[obq + c], k = par[obk + c], v = par[obv + c], s = par[obs + c];
#pragma unroll
    for (int d = 0; d < TF_H; ++d) {
      const float x = p.X[i * TF_H + d];
      q = fmaf(x, p.WqT[d * 16 + c], q); k = fmaf(x, p.WkT[d * 16 + c], k);
      v = fmaf(x, p.WvT[d * 16 + c], v); s = fmaf(x, p.WsT[d * 16 + c], s);
    }
    p.Q[idx] = q; p.K[idx] = k; p.V[idx] = v; p.B2[idx] = s;
  }
  // ---- edge MLP hidden layer
  for (int idx = tid; idx < E * TF_K; idx += kTfThreads) {
    const int e = idx >> 3, k = idx & 7;
    float h = par[ob1 + k];
#pragma unroll
    for (int d = 0; d < TF_D; ++d) h = fmaf(p.attr[e * TF_D + d], par[oW1 + k * TF_D + d], h);
    p.hid[idx] = fmaxf(h, 0.f);
  }
  __syncthreads();
  TF_STAMP(3);
  // ---- TransformerConv attention: 16 lanes (channels) per destination node, in-edges in edge order
  const int cl = tid & 15;                              // this thread's channel in every (node, channel) loop
  const float we0 = par[oWe + cl * TF_D], we1 = par[oWe + cl * TF_D + 1], we2 = par[oWe + cl * TF_D + 2],
              we3 = par[oWe + cl * TF_D + 3];
  for (int base = 0; base < n * TF_H; base += kTfThreads) {
    const int idx = base + tid, i = idx >> 4, c = cl;
    const bool valid = idx < n * TF_H;
    const int r0 = valid ? p.rowd[i] : 0, r1 = valid ? p.rowd[i + 1] : 0;
    int rmax = r1 - r0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = max(rmax, __shfl_xor_sync(kFull, rmax, o));
    const float qi = valid ? p.Q[idx] : 0.f;
    float mx = -INFINITY;
    for (int r = 0; r < rmax; ++r) {                    // pass 1: logits (kept in alpha[]) and their maximum
      const bool on = r0 + r < r1;
      const int e = on ? p.permd[r0 + r] : 0;
      const float* at = p.attr + e * TF_D;
      const float ee = fmaf(at[3], we3, fmaf(at[2], we2, fmaf(at[1], we1, at[0] * we0)));
      const float lg = tf_sum16(on ? qi * (p.K[p.src[e] * TF_H + c] + ee) : 0.f) * 0.25f;
      if (on) {
        mx = fmaxf(mx, lg);
        if (c == 0) p.alpha[e] = lg;
      }
    }
    __syncwarp();
    float den = 0.f, acc = 0.f;
    for (int r = r0; r < r1; ++r) {                     // pass 2: one exp per edge; unnormalised weights back into alpha[]
      const int e = p.permd[r];
      const float* at = p.attr + e * TF_D;
      const float ee = fmaf(at[3], we3, fmaf(at[2], we2, fmaf(at[1], we1, at[0] * we0)));
      const float w = expf(p.alpha[e] - mx);
      den += w;
      acc = fmaf(w, p.V[p.src[e] * TF_H + c] + ee, acc);
    }
    const float inv = 1.f / (den + 1e-16f);
    __syncwarp();
    if (c == 0)
      for (int r = r0; r < r1; ++r) { const int e = p.permd[r]; p.alpha[e] = expf(p.alpha[e] - mx) * inv; }
    if (valid) p.H1[idx] = tf_lk(fmaf(acc, inv, p.B2[idx])) * tf_keep(dr.m1, idx, dr.scale);
  }
  __syncthreads();
  TF_STAMP(4);
  // ---- factorised NNConv: T_j = h1_j P
  for (int idx = tid; idx < n * (TF_T / 4); idx += kTfThreads) {
    const int j = idx / (TF_T / 4), r = 4 * (idx % (TF_T / 4));
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < TF_H; ++c) {
      const float h = p.H1[j * TF_H + c];
      const float4 w = __ldg(reinterpret_cast<const float4*>(par + oW2 + c * TF_T + r));
      t.x = fmaf(h, w.x, t.x); t.y = fmaf(h, w.y, t.y); t.z = fmaf(h, w.z, t.z); t.w = fmaf(h, w.w, t.w);
    }
    *reinterpret_cast<float4*>(p.T + j * TF_T + r) = t;
  }
  __syncthreads();
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int i = idx >> 4, o = idx & 15;
    const int r0 = p.rowd[i], r1 = p.rowd[i + 1];
    float m = 0.f;
    for (int r = r0; r < r1; ++r) {
      const int e = p.permd[r];
      const float* t = p.T + p.src[e] * TF_T + o;
      float me = t[TF_K * TF_H];                        // the bias slab (hid' = 1)
#pragma unroll
      for (int k = 0; k < TF_K; ++k) me = fmaf(p.hid[e * TF_K + k], t[k * TF_H], me);
      m += me;
    }
    float v = m / static_cast<float>(max(r1 - r0, 1)) + par[obias2 + o];
#pragma unroll
    for (int c = 0; c < TF_H; ++c) v = fmaf(p.H1[i * TF_H + c], p.WrootT[c * 16 + o], v);
    p.B1[idx] = tf_lk(v) * tf_keep(dr.m2, idx, dr.scale);
  }
  __syncthreads();
  TF_STAMP(5);
  // ---- global mean pool + MLP head (one warp)
  float* pool = p.small; float* pre1 = p.small + 16; float* z1 = p.small + 32; float* outv = p.small + 48;
  if (tid < TF_H) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += p.B1[i * TF_H + tid];
    pool[tid] = s / static_cast<float>(max(n, 1));
  }
  __syncthreads();
  if (tid < TF_H) {
    float h = par[obm1 + tid];
#pragma unroll
    for (int o = 0; o < TF_H; ++o) h = fmaf(pool[o], p.Wm1T[o * 16 + tid], h);
    pre1[tid] = h;
    z1[tid] = tf_lk(h) * tf_keep(dr.m3, tid, dr.scale_h);
  }
  __syncthreads();
  if (tid < QOT_OUT) {
    float o3 = par[obm2 + tid];
#pragma unroll
    for (int o = 0; o < TF_H; ++o) o3 = fmaf(z1[o], par[oWm2 + tid * TF_H + o], o3);
    outv[tid] = o3;
  }
  __syncthreads();
  if (save_n) {
    const int nh = n * TF_H;
    for (int idx = tid; idx < nh; idx += kTfThreads) {
      save_n[idx] = p.Q[idx]; save_n[nh + idx] = p.K[idx]; save_n[2 * nh + idx] = p.V[idx];
      save_n[3 * nh + idx] = p.H1[idx]; save_n[4 * nh + idx] = p.B1[idx];
    }
    for (int idx = tid; idx < n * TF_T; idx += kTfThreads) save_n[5 * nh + idx] = p.T[idx];
    for (int idx = tid; idx < E * TF_K; idx += kTfThreads) save_e[idx] = p.hid[idx];
    for (int idx = tid; idx < E; idx += kTfThreads) save_e[E * TF_K + idx] = p.alpha[idx];
    if (tid < 52) save_g[tid] = p.small[tid];
  }
}

__global__ void __launch_bounds__(kTfThreads, 7)
topo_fused_fwd_kernel(const float* __restrict__ prep, const float* __restrict__ emb, const int64_t* __restrict__ node_ids,
                      const int64_t* __restrict__ edge_index, int64_t Etot, const float* __restrict__ eattr,
                      const int64_t* __restrict__ gptr, const int64_t* __restrict__ eptr, int64_t B, int nmax, int emax,
                      int num_nodes, float* __restrict__ out, float* __restrict__ saved, int64_t Ntot,
                      int32_t* __restrict__ status, const uint8_t* __restrict__ drop_mask, float drop_scale,
                      float drop_scale_head) {
  extern __shared__ __align__(16) char tf_smem[];
  const TfPtrs p = tf_carve(tf_smem, prep, nmax, emax, num_nodes, false);
  float* sv_e = saved ? saved + Ntot * kTfSaveNode : nullptr;
  float* sv_g = saved ? sv_e + Etot * kTfSaveEdge : nullptr;
  for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
    const int64_t n0 = gptr[g], e0 = eptr[g];
    const int64_t n = gptr[g + 1] - n0, E = eptr[g + 1] - e0;
    if (n < 0 || E < 0 || n > nmax || E > emax) {       // caps come from the same arrays on the host: never expected
      if (threadIdx.x == 0) atomicOr(status, 1);
      if (threadIdx.x < QOT_OUT) out[g * QOT_OUT + threadIdx.x] = __int_as_float(0x7fc00000);   // NaN, never stale memory
      continue;
    }
    tf_graph_forward(p, emb, node_ids, edge_index, edge_index + Etot, eattr, n0, static_cast<int>(n), e0, static_cast<int>(E), num_nodes,
                     saved ? saved + n0 * kTfSaveNode : nullptr, saved ? sv_e + e0 * kTfSaveEdge : nullptr,
                     saved ? sv_g + g * kTfSaveGraph : nullptr, nullptr, nullptr, nullptr, status,
                     tf_drop_of(drop_mask, drop_scale, drop_scale_head, Ntot, n0, g));
    if (threadIdx.x < QOT_OUT) out[g * QOT_OUT + threadIdx.x] = p.small[48 + threadIdx.x];
    __syncthreads();
  }
}

// forward recomputed, then the backward of one graph given d loss / d out; gradients accumulate in p.grad
// (flat layout, P layout for W2 | b2) and p.gemb
__device__ void tf_graph_backward(const TfPtrs& p, const int64_t* __restrict__ node_ids, int64_t n0, int n, int E,
                                  int num_nodes, const TfDrop& dr) {
  const int tid = threadIdx.x;
  const float* par = p.par;
  float* g = p.grad;
  float* pool = p.small; float* pre1 = p.small + 16; float* z1 = p.small + 32;
  float* dpool = p.small + 64; float* dpre1 = p.small + 80; const float* dout = p.small + 52;
  TF_STAMP(6);
  // ---- MLP head
  if (tid < TF_H) {
    float dz = 0.f;
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) dz = fmaf(par[oWm2 + k * TF_H + tid], dout[k], dz);
    dpre1[tid] = dz * tf_keep(dr.m3, tid, dr.scale_h) * (pre1[tid] > 0.f ? 1.f : kTfSlope);
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) g[oWm2 + k * TF_H + tid] += dout[k] * z1[tid];
    if (tid < QOT_OUT) g[obm2 + tid] += dout[tid];
  }
  __syncthreads();
  for (int idx = tid; idx < 256; idx += kTfThreads) g[oWm1 + idx] += dpre1[idx >> 4] * pool[idx & 15];
  if (tid < TF_H) {
    g[obm1 + tid] += dpre1[tid];
    float d = 0.f;
#pragma unroll
    for (int o = 0; o < TF_H; ++o) d = fmaf(par[oWm1 + o * TF_H + tid], dpre1[o], d);
    dpool[tid] = d / static_cast<float>(max(n, 1));
  }
  __syncthreads();
  TF_STAMP(7);
  // ---- dO2 (in place of H2 in B1); dH1 starts in B2
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads)
    p.B1[idx] = dpool[idx & 15] * tf_keep(dr.m2, idx, dr.scale) * tf_dlk_from_act(p.B1[idx]);
  __syncthreads();
  if (tid < TF_H) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += p.B1[i * TF_H + tid];
    g[obias2 + tid] += s;
  }
  for (int idx = tid; idx < 64; idx += kTfThreads) {    // dWroot[o][c..c+3] += sum_i dO2[i][o] h1[i][c..c+3]
    const int o = idx >> 2, c = 4 * (idx & 3);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
      const float d = p.B1[i * TF_H + o];
      const float4 h = *reinterpret_cast<const float4*>(p.H1 + i * TF_H + c);
      acc.x = fmaf(d, h.x, acc.x); acc.y = fmaf(d, h.y, acc.y); acc.z = fmaf(d, h.z, acc.z); acc.w = fmaf(d, h.w, acc.w);
    }
    float* gp = g + oWroot + o * TF_H + c;
    gp[0] += acc.x; gp[1] += acc.y; gp[2] += acc.z; gp[3] += acc.w;
  }
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int i = idx >> 4, c = idx & 15;
    float d = 0.f;
#pragma unroll
    for (int o = 0; o < TF_H; ++o) d = fmaf(par[oWroot + o * TF_H + c], p.B1[i * TF_H + o], d);
    p.B2[idx] = d;
  }
  TF_STAMP(8);
  // ---- dm_i = dO2_i / deg_i (G3: free until the attention backward)
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int i = idx >> 4;
    p.G3[idx] = p.B1[idx] / static_cast<float>(max(p.rowd[i + 1] - p.rowd[i], 1));
  }
  __syncthreads();
  // ---- d hid' per edge (needs T of the source), edge-MLP layer-1 gradients
  for (int idx = tid; idx < E * TF_K; idx += kTfThreads) {
    const int e = idx >> 3, k = idx & 7, i = p.dst[e];
    const float* t = p.T + p.src[e] * TF_T + k * TF_H;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int o = 0; o < TF_H; o += 2) {
      d0 = fmaf(t[o], p.G3[i * TF_H + o], d0);
      d1 = fmaf(t[o + 1], p.G3[i * TF_H + o + 1], d1);
    }
    p.dhid[idx] = p.hid[idx] > 0.f ? d0 + d1 : 0.f;
  }
  __syncthreads();
  if (tid < TF_K * TF_D) {                              // dW1[k][d] += sum_e dhidpre[e][k] attr[e][d]
    const int k = tid >> 2, d = tid & 3;
    float s = 0.f;
    for (int e = 0; e < E; ++e) s = fmaf(p.dhid[e * TF_K + k], p.attr[e * TF_D + d], s);
    g[oW1 + tid] += s;
  } else if (tid < TF_K * TF_D + TF_K) {
    const int k = tid - TF_K * TF_D;
    float s = 0.f;
    for (int e = 0; e < E; ++e) s += p.dhid[e * TF_K + k];
    g[ob1 + k] += s;
  }
  __syncthreads();
  TF_STAMP(9);
  // ---- dT_j (over the out-edges of j, edge order) overwrites T_j
  for (int idx = tid; idx < n * (TF_T / 4); idx += kTfThreads) {
    const int j = idx / (TF_T / 4), r = 4 * (idx % (TF_T / 4)), k = r >> 4, o = r & 15;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = p.rows[j]; q < p.rows[j + 1]; ++q) {
      const int e = p.perms[q], i = p.dst[e];
      const float hk = k < TF_K ? p.hid[e * TF_K + k] : 1.f;
      const float4 dm = *reinterpret_cast<const float4*>(p.G3 + i * TF_H + o);
      acc.x = fmaf(hk, dm.x, acc.x); acc.y = fmaf(hk, dm.y, acc.y); acc.z = fmaf(hk, dm.z, acc.z); acc.w = fmaf(hk, dm.w, acc.w);
    }
    *reinterpret_cast<float4*>(p.T + j * TF_T + r) = acc;
  }
  __syncthreads();
  TF_STAMP(10);
  for (int idx = tid; idx < TF_H * (TF_T / 4); idx += kTfThreads) {   // dP[c][r..r+3] += sum_j h1[j][c] dT_j[r..r+3]
    const int c = idx / (TF_T / 4), r = 4 * (idx % (TF_T / 4));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      const float h = p.H1[j * TF_H + c];
      const float4 t = *reinterpret_cast<const float4*>(p.T + j * TF_T + r);
      acc.x = fmaf(h, t.x, acc.x); acc.y = fmaf(h, t.y, acc.y); acc.z = fmaf(h, t.z, acc.z); acc.w = fmaf(h, t.w, acc.w);
    }
    float4* gp = reinterpret_cast<float4*>(g + oW2 + c * TF_T + r);
    float4 cur = *gp;
    cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
    *gp = cur;
  }
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {      // dH1[j][c] += sum_r P[c][r] dT_j[r]; then dO1
    const int j = idx >> 4, c = idx & 15;
    const float4* pr = reinterpret_cast<const float4*>(par + oW2 + c * TF_T);
    const float4* t = reinterpret_cast<const float4*>(p.T + j * TF_T);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 4
    for (int r = 0; r < TF_T / 4; ++r) {
      const float4 w = __ldg(pr + r), v = t[r];
      d0 = fmaf(w.x, v.x, d0); d1 = fmaf(w.y, v.y, d1); d2 = fmaf(w.z, v.z, d2); d3 = fmaf(w.w, v.w, d3);
    }
    p.B2[idx] = (p.B2[idx] + ((d0 + d1) + (d2 + d3))) * tf_keep(dr.m1, idx, dr.scale) * tf_dlk_from_act(p.H1[idx]);
  }
  __syncthreads();
  TF_STAMP(11);
  // ---- skip projection
  if (tid < TF_H) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += p.B2[i * TF_H + tid];
    g[obs + tid] += s;
  }
  for (int idx = tid; idx < 64; idx += kTfThreads) {
    const int c = idx >> 2, k = 4 * (idx & 3);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
      const float d = p.B2[i * TF_H + c];
      const float4 xv = *reinterpret_cast<const float4*>(p.X + i * TF_H + k);
      acc.x = fmaf(d, xv.x, acc.x); acc.y = fmaf(d, xv.y, acc.y); acc.z = fmaf(d, xv.z, acc.z); acc.w = fmaf(d, xv.w, acc.w);
    }
    float* gp = g + oWs + c * TF_H + k;
    gp[0] += acc.x; gp[1] += acc.y; gp[2] += acc.z; gp[3] += acc.w;
  }
  TF_STAMP(12);
  // ---- attention backward, destination side: dalpha, dlogit (per edge), dQ (G1)
  __syncthreads();
  const int cb = tid & 15;
  const float we0 = par[oWe + cb * TF_D], we1 = par[oWe + cb * TF_D + 1], we2 = par[oWe + cb * TF_D + 2],
              we3 = par[oWe + cb * TF_D + 3];
  for (int base = 0; base < n * TF_H; base += kTfThreads) {
    const int idx = base + tid, i = idx >> 4, c = idx & 15;
    const bool valid = idx < n * TF_H;
    const int r0 = valid ? p.rowd[i] : 0, r1 = valid ? p.rowd[i + 1] : 0;
    int rmax = r1 - r0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = max(rmax, __shfl_xor_sync(kFull, rmax, o));
    const float gi = valid ? p.B2[idx] : 0.f, qi = valid ? p.Q[idx] : 0.f;
    float tsum = 0.f;
    for (int r = 0; r < rmax; ++r) {                    // pass 1: dalpha_e (kept in dlogit[]), t = sum alpha dalpha
      const bool on = r0 + r < r1;
      const int e = on ? p.permd[r0 + r] : 0;
      const float* at = p.attr + e * TF_D;
      const float ee = fmaf(at[3], we3, fmaf(at[2], we2, fmaf(at[1], we1, at[0] * we0)));
      const float da = tf_sum16(on ? gi * (p.V[p.src[e] * TF_H + c] + ee) : 0.f);
      if (on) {
        tsum = fmaf(p.alpha[e], da, tsum);
        if (c == 0) p.dlogit[e] = da;
      }
    }
    __syncwarp();
    float dq = 0.f;
    for (int r = r0; r < r1; ++r) {
      const int e = p.permd[r];
      const float* at = p.attr + e * TF_D;
      const float ee = fmaf(at[3], we3, fmaf(at[2], we2, fmaf(at[1], we1, at[0] * we0)));
      const float dl = p.alpha[e] * (p.dlogit[e] - tsum);
      dq = fmaf(dl, (p.K[p.src[e] * TF_H + c] + ee) * 0.25f, dq);
    }
    __syncwarp();
    for (int r = r0; r < r1; ++r) {
      const int e = p.permd[r];
      if (c == 0) p.dlogit[e] = p.alpha[e] * (p.dlogit[e] - tsum);
    }
    if (valid) p.G1[idx] = dq;
    (void)qi;
  }
  __syncthreads();
  TF_STAMP(13);
  // ---- source side: dV (G3), dK (G2) over the out-edges of j, edge order
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {
    const int j = idx >> 4, c = idx & 15;
    float dv = 0.f, dk = 0.f;
    for (int q = p.rows[j]; q < p.rows[j + 1]; ++q) {
      const int e = p.perms[q], i = p.dst[e];
      dv = fmaf(p.alpha[e], p.B2[i * TF_H + c], dv);
      dk = fmaf(p.dlogit[e], p.Q[i * TF_H + c] * 0.25f, dk);
    }
    p.G3[idx] = dv; p.G2[idx] = dk;
  }
  if (tid < TF_H * TF_D) {                                // dWe[c][d] += sum_e (dmsg_e[c] + dkey_e[c]) attr[e][d]
    const int c = tid >> 2, d = tid & 3;
    float s = 0.f;
    for (int e = 0; e < E; ++e) {
      const int i = p.dst[e];
      s = fmaf(fmaf(p.alpha[e], p.B2[i * TF_H + c], p.dlogit[e] * p.Q[i * TF_H + c] * 0.25f), p.attr[e * TF_D + d], s);
    }
    g[oWe + tid] += s;
  }
  __syncthreads();
  TF_STAMP(14);
  // ---- q / k / v weight gradients, dX, embedding rows
  if (tid < 3 * TF_H) {
    const int w = tid >> 4, c = tid & 15;
    const float* G = w == 0 ? p.G1 : (w == 1 ? p.G2 : p.G3);
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += G[i * TF_H + c];
    g[(w == 0 ? obq : (w == 1 ? obk : obv)) + c] += s;
  }
  for (int idx = tid; idx < 3 * 64; idx += kTfThreads) {
    const int w = idx >> 6, c = (idx >> 2) & 15, k = 4 * (idx & 3);
    const float* G = w == 0 ? p.G1 : (w == 1 ? p.G2 : p.G3);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
      const float d = G[i * TF_H + c];
      const float4 xv = *reinterpret_cast<const float4*>(p.X + i * TF_H + k);
      acc.x = fmaf(d, xv.x, acc.x); acc.y = fmaf(d, xv.y, acc.y); acc.z = fmaf(d, xv.z, acc.z); acc.w = fmaf(d, xv.w, acc.w);
    }
    float* gp = g + (w == 0 ? oWq : (w == 1 ? oWk : oWv)) + c * TF_H + k;
    gp[0] += acc.x; gp[1] += acc.y; gp[2] += acc.z; gp[3] += acc.w;
  }
  for (int idx = tid; idx < n * TF_H; idx += kTfThreads) {      // dX[i][k] into B1 (dO2 is dead)
    const int i = idx >> 4, k = idx & 15;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < TF_H; ++c) {
      d0 = fmaf(p.B2[i * TF_H + c], par[oWs + c * TF_H + k], d0);
      d1 = fmaf(p.G1[i * TF_H + c], par[oWq + c * TF_H + k], d1);
      d2 = fmaf(p.G2[i * TF_H + c], par[oWk + c * TF_H + k], d2);
      d3 = fmaf(p.G3[i * TF_H + c], par[oWv + c * TF_H + k], d3);
    }
    p.T[idx] = (d0 + d1) + (d2 + d3);                     // T is dead: dX lives there
  }
  __syncthreads();
  TF_STAMP(15);
  if (tid < TF_H) {                                       // one thread per column, nodes in order: duplicates of an id are summed deterministically
    for (int i = 0; i < n; ++i) {
      long long id = node_ids[n0 + i];
      id = id < 0 ? 0 : (id >= num_nodes ? num_nodes - 1 : id);
      p.gemb[id * TF_H + tid] += p.T[i * TF_H + tid];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kTfThreads, 7)
topo_fused_bwd_kernel(const float* __restrict__ prep, const float* __restrict__ emb, const int64_t* __restrict__ node_ids,
                      const int64_t* __restrict__ edge_index, int64_t Etot, const float* __restrict__ eattr,
                      const int64_t* __restrict__ gptr, const int64_t* __restrict__ eptr, int64_t B, int nmax, int emax,
                      int num_nodes, const float* __restrict__ dout, const float* __restrict__ saved, int64_t Ntot,
                      float* __restrict__ partial, int32_t* __restrict__ status, const uint8_t* __restrict__ drop_mask,
                      float drop_scale, float drop_scale_head) {
  extern __shared__ __align__(16) char tf_smem[];
  const TfPtrs p = tf_carve(tf_smem, prep, nmax, emax, num_nodes, true);
  const float* sv_e = saved ? saved + Ntot * kTfSaveNode : nullptr;
  const float* sv_g = saved ? sv_e + Etot * kTfSaveEdge : nullptr;
  const int gsz = kTfParams + num_nodes * TF_H;
  for (int i = threadIdx.x; i < gsz; i += kTfThreads) p.grad[i] = 0.f;       // grad | gemb are contiguous
  __syncthreads();
  for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
    const int64_t n0 = gptr[g], e0 = eptr[g];
    const int64_t n = gptr[g + 1] - n0, E = eptr[g + 1] - e0;
    if (n < 0 || E < 0 || n > nmax || E > emax) {
      if (threadIdx.x == 0) atomicOr(status, 1);
      continue;
    }
    TF_STAMP(0);
    tf_graph_forward(p, emb, node_ids, edge_index, edge_index + Etot, eattr, n0, static_cast<int>(n), e0, static_cast<int>(E), num_nodes,
                     nullptr, nullptr, nullptr, saved ? saved + n0 * kTfSaveNode : nullptr,
                     saved ? sv_e + e0 * kTfSaveEdge : nullptr, saved ? sv_g + g * kTfSaveGraph : nullptr, status,
                     tf_drop_of(drop_mask, drop_scale, drop_scale_head, Ntot, n0, g));
    if (threadIdx.x < QOT_OUT) p.small[52 + threadIdx.x] = dout[g * QOT_OUT + threadIdx.x];
    __syncthreads();
    tf_graph_backward(p, node_ids, n0, static_cast<int>(n), static_cast<int>(E), num_nodes,
                      tf_drop_of(drop_mask, drop_scale, drop_scale_head, Ntot, n0, g));
    TF_STAMP(16);
  }
  float* dst = partial + static_cast<size_t>(blockIdx.x) * gsz;
  for (int i = threadIdx.x; i < gsz; i += kTfThreads) dst[i] = p.grad[i];
}

// per-block partials -> flat gradient; W2 | b2 back from the P layout.  256 threads = 8 slices of the block
// range x 32 gradient elements; slice sums are combined in slice order: a fixed summation tree
__global__ void __launch_bounds__(256)
topo_fused_reduce_kernel(const float* __restrict__ partial, int blocks, int num_nodes,
                         float* __restrict__ gflat, float* __restrict__ gemb) {
  __shared__ float part[8][32];
  const int gsz = kTfParams + num_nodes * TF_H;
  const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  float s = 0.f;
  if (i < gsz)
    for (int b = sl; b < blocks; b += 8) s += partial[static_cast<size_t>(b) * gsz + i];
  part[sl][col] = s;
  __syncthreads();
  if (sl != 0 || i >= gsz) return;
#pragma unroll
  for (int r = 1; r < 8; ++r) s += part[r][col];
  if (i >= kTfParams) { gemb[i - kTfParams] = s; return; }
  if (i >= oW2 && i < oWroot) {
    const int q = i - oW2, c = q / TF_T, r = q % TF_T, k = r / TF_H, o = r % TF_H;
    if (k < TF_K) gflat[oW2 + (c * TF_H + o) * TF_K + k] = s; else gflat[ob2 + c * TF_H + o] = s;
    return;
  }
  gflat[i] = s;
}

}  // namespace qot

using namespace qot;

#ifdef QOT_LP_TRACE
extern "C" int qot_debug_set_tf_trace(long long* buf) { return cudaMemcpyToSymbol(g_tf_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -3; }
#endif
extern "C" int qot_topo_fused_params(void) { return kTfParams; }
extern "C" int qot_topo_fused_prepared_floats(void) { return kTfPrepared; }

extern "C" int qot_topo_fused_prepare(const float* flat, float* prepared, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(flat && prepared, "qot_topo_fused_prepare: null argument");
  topo_fused_prepare_kernel<<<16, 256, 0, stream>>>(flat, prepared);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

static int tf_blocks(int64_t B, size_t smem) {
  // resident blocks per SM: shared memory, 64 registers x 128 threads, at most 8
  const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(7, (228 * 1024) / (smem + 1024))));
  return static_cast<int>(std::min<int64_t>(B, static_cast<int64_t>(per_sm) * kNumSMs));
}

static int tf_check(int64_t B, int nmax, int emax, int num_nodes, const char* who) {
  QOT_REQUIRE(B >= 0 && nmax >= 0 && emax >= 0 && num_nodes > 0, "%s: bad sizes", who);
  QOT_REQUIRE(nmax <= 4096 && emax <= 4 * kTfThreads, "%s: graph too large for one block (at most %d edges)", who, 4 * kTfThreads);
  return QOT_OK;
}

extern "C" size_t qot_topo_fused_saved_floats(int64_t N, int64_t E, int64_t B) {
  return static_cast<size_t>(std::max<int64_t>(N, 0)) * kTfSaveNode + static_cast<size_t>(std::max<int64_t>(E, 0)) * kTfSaveEdge +
         static_cast<size_t>(std::max<int64_t>(B, 0)) * kTfSaveGraph + 4;
}

extern "C" int qot_topo_fused_fwd(const float* prepared, const float* emb, const int64_t* node_ids, const int64_t* edge_index,
                                  int64_t Etot, const float* edge_attr, const int64_t* gptr, const int64_t* eptr,
                                  int64_t B, int64_t N, int32_t nmax, int32_t emax, int32_t num_nodes, float* out,
                                  float* saved, int32_t* status, const uint8_t* drop_mask, float drop_scale,
                                  float drop_scale_head, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = tf_check(B, nmax, emax, num_nodes, "qot_topo_fused_fwd");
  if (rc) return rc;
  QOT_REQUIRE(prepared && emb && gptr && eptr && status && (B == 0 || out), "qot_topo_fused_fwd: null argument");
  if (B == 0) return QOT_OK;
  const size_t smem = tf_smem_bytes(nmax, emax, num_nodes, false);
  QOT_REQUIRE(smem <= 227 * 1024, "qot_topo_fused_fwd: graphs of %d nodes / %d edges need %zu bytes of shared memory", nmax, emax, smem);
  QOT_CUDA(cudaFuncSetAttribute(topo_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int blocks = tf_blocks(B, smem);
  topo_fused_fwd_kernel<<<blocks, kTfThreads, smem, stream>>>(prepared, emb, node_ids, edge_index, Etot, edge_attr, gptr, eptr, B,
                                                             nmax, emax, num_nodes, out, saved, N, status, drop_mask, drop_scale,
                                                             drop_scale_head);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_topo_fused_bwd_workspace_bytes(int32_t num_nodes) {
  return static_cast<size_t>(8 * kNumSMs) * (kTfParams + static_cast<size_t>(std::max(num_nodes, 0)) * TF_H) * 4 + 256;
}

extern "C" int qot_topo_fused_bwd(const float* prepared, const float* emb, const int64_t* node_ids, const int64_t* edge_index,
                                  int64_t Etot, const float* edge_attr, const int64_t* gptr, const int64_t* eptr,
                                  int64_t B, int64_t N, int32_t nmax, int32_t emax, int32_t num_nodes, const float* dout,
                                  const float* saved, float* gflat, float* gemb, void* ws, size_t ws_bytes,
                                  int32_t* status, const uint8_t* drop_mask, float drop_scale, float drop_scale_head,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = tf_check(B, nmax, emax, num_nodes, "qot_topo_fused_bwd");
  if (rc) return rc;
  QOT_REQUIRE(prepared && emb && gptr && eptr && status && gflat && gemb && (B == 0 || dout), "qot_topo_fused_bwd: null argument");
  QOT_REQUIRE(ws && ws_bytes >= qot_topo_fused_bwd_workspace_bytes(num_nodes), "qot_topo_fused_bwd: workspace too small");
  const int gsz = kTfParams + num_nodes * TF_H;
  if (B == 0) {
    QOT_CUDA(cudaMemsetAsync(gflat, 0, kTfParams * 4, stream));
    QOT_CUDA(cudaMemsetAsync(gemb, 0, static_cast<size_t>(num_nodes) * TF_H * 4, stream));
    return QOT_OK;
  }
  const size_t smem = tf_smem_bytes(nmax, emax, num_nodes, true);
  QOT_REQUIRE(smem <= 227 * 1024, "qot_topo_fused_bwd: graphs of %d nodes / %d edges need %zu bytes of shared memory", nmax, emax, smem);
  QOT_CUDA(cudaFuncSetAttribute(topo_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int blocks = tf_blocks(B, smem);
  topo_fused_bwd_kernel<<<blocks, kTfThreads, smem, stream>>>(prepared, emb, node_ids, edge_index, Etot, edge_attr, gptr, eptr, B,
                                                             nmax, emax, num_nodes, dout, saved, N, static_cast<float*>(ws), status,
                                                             drop_mask, drop_scale, drop_scale_head);
  QOT_LAUNCH_CHECK();
  topo_fused_reduce_kernel<<<(gsz + 31) / 32, 256, 0, stream>>>(static_cast<const float*>(ws), blocks, num_nodes, gflat, gemb);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
