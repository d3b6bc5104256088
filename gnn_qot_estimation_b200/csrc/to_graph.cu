// Lightpath graph construction on the device: to_graph.py::create_lightpath_graph (:187-312) fused
// with the tensorisation of lightpath_training/dataset.py:53-123.  One 256-thread block per sample;
// everything between the raw [lp_feat, link, freq] tensor and the packed graph store stays in shared
// memory:
//   1. row-major scan of the (link, freq) channels, occupied ones compacted IN ORDER (block scan);
//   2. conn_id -> first occupied channel through a shared-memory hash table (atomicCAS claims a slot,
//      atomicMin keeps the earliest channel), so node order = order of first appearance;
//   3. node rows from the first channel (fp64 min-max scaling, then fp32 -- bit-exact with the
//      Python floats of dataset.py:77-80);
//   4. the 0 < |df| < threshold join per link over the link's contiguous run of entries, fp64 like
//      numpy (to_graph.py:296-299), links used by one lightpath skipped (:285), edges as bits of an
//      n x n adjacency matrix (set semantics of nx.Graph.add_edge for free);
//   5. directed edges written sorted by (source, target) from the bit rows.
// Integer / index work throughout: HBM-bound on the one read of the sample tensor (F*L*Q*4 bytes).
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int kTgThreads = 256;
constexpr int kTgHash = 1024;                 // slots (>= 4 x QOT_TG_MAX_NODES), open addressing
constexpr int kTgAdjWords = QOT_TG_MAX_NODES / 32;

struct TgSmem {
  int conn[QOT_TG_MAX_CHANNELS];              // conn_id of the k-th occupied channel (row-major order)
  unsigned short link[QOT_TG_MAX_CHANNELS], freq[QOT_TG_MAX_CHANNELS], node[QOT_TG_MAX_CHANNELS];
  int hkey[kTgHash], hmin[kTgHash], hnode[kTgHash];
  int seg[QOT_TG_MAX_LINKS + 1];              // first entry of every link's run
  unsigned adj[QOT_TG_MAX_NODES][kTgAdjWords];
  int first[QOT_TG_MAX_NODES];                // entry index of each node's first channel
  int scan[kTgThreads / 32];
  int total, bad;
};

// exclusive prefix of `v` over the block (256 threads); *total = block sum.  Two barriers.
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += up;
  }
  __syncthreads();                              // previous use of warp_sums is over
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kTgThreads / 32; ++w) {
    const int s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + incl - v;
}
__device__ __forceinline__ unsigned tg_hash(int c) { return (static_cast<unsigned>(c) * 2654435761u) >> 22; }

// Steps 1-2, shared by both representations: occupied channels compacted in row-major order, conn_id ->
// earliest entry, nodes ranked by first appearance.  Returns K (entries), n (distinct conn_ids), bad.
__device__ __forceinline__ void tg_scan_and_rank(TgSmem& sm, const float* __restrict__ d, const qot_lp_graph_cfg_t& cfg,
                                                 int& K_out, int& n_out, bool& bad_out) {
  const int tid = threadIdx.x;
  const int L = cfg.L, Q = cfg.Q, LQ = L * Q;
  for (int i = tid; i < kTgHash; i += kTgThreads) { sm.hkey[i] = INT_MIN; sm.hmin[i] = INT_MAX; sm.hnode[i] = -1; }
  for (int i = tid; i < QOT_TG_MAX_NODES * kTgAdjWords; i += kTgThreads) (&sm.adj[0][0])[i] = 0u;
  for (int i = tid; i <= min(L, QOT_TG_MAX_LINKS); i += kTgThreads) sm.seg[i] = -1;
  if (tid == 0) sm.bad = 0;
  __syncthreads();

  // ---- 1. occupied channels, compacted in row-major order (to_graph.py:229-232).  A thread takes 4
  // consecutive channels per round (one 16-byte load per lp_feat row, all rows in flight at once), the
  // block one prefix sum per 1024 channels
  int K = 0;
  const bool vec = (LQ & 3) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0;
  for (int c0 = 0; c0 < LQ; c0 += 4 * kTgThreads) {
    const int ch0 = c0 + 4 * tid;
    unsigned occ4 = 0u;
    float cv[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec && ch0 < LQ) {                              // LQ % 4 == 0: the quad is entirely inside
#pragma unroll 5
      for (int f = 0; f < cfg.F; ++f) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(d + static_cast<int64_t>(f) * LQ + ch0));
        occ4 |= (v.x != 0.f ? 1u : 0u) | (v.y != 0.f ? 2u : 0u) | (v.z != 0.f ? 4u : 0u) | (v.w != 0.f ? 8u : 0u);
        if (f == cfg.i_conn) { cv[0] = v.x; cv[1] = v.y; cv[2] = v.z; cv[3] = v.w; }
      }
    } else if (!vec) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ch0 + u < LQ) {
          for (int f = 0; f < cfg.F; ++f) occ4 |= (d[static_cast<int64_t>(f) * LQ + ch0 + u] != 0.f ? 1u : 0u) << u;
          cv[u] = d[static_cast<int64_t>(cfg.i_conn) * LQ + ch0 + u];
        }
      }
    }
    int tot;
    int pos = K + block_excl_scan(__popc(occ4), sm.scan, &tot);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if ((occ4 >> u) & 1u) {
        if (pos < QOT_TG_MAX_CHANNELS) {
          sm.conn[pos] = static_cast<int>(cv[u]);                            // int(): toward zero (:243)
          sm.link[pos] = static_cast<unsigned short>((ch0 + u) / Q);
          sm.freq[pos] = static_cast<unsigned short>((ch0 + u) % Q);
        } else {
          sm.bad = 1;
        }
        ++pos;
      }
    }
    K += tot;
  }
  __syncthreads();
  bool bad = sm.bad != 0 || L > QOT_TG_MAX_LINKS || K > QOT_TG_MAX_CHANNELS;
  if (bad) K = 0;

  // ---- 2. conn_id -> earliest entry (dict insertion order of to_graph.py:245-268)
  for (int k = tid; k < K; k += kTgThreads) {
    const int c = sm.conn[k];
    unsigned h = tg_hash(c);
    bool placed = false;
    for (int probe = 0; probe < kTgHash && !placed; ++probe, h = (h + 1) & (kTgHash - 1)) {
      const int prev = atomicCAS(&sm.hkey[h], INT_MIN, c);
      if (prev == INT_MIN || prev == c) {
        atomicMin(&sm.hmin[h], k);
        placed = true;
      }
    }
    if (!placed || c == INT_MIN) sm.bad = 1;           // more distinct conn_ids than slots: over capacity
  }
  __syncthreads();
  if (sm.bad) { bad = true; K = 0; }                   // (also keeps every lookup below finite)
  // nodes in order of first appearance: rank of the "first" entries
  int n = 0;
  for (int k0 = 0; k0 < K; k0 += kTgThreads) {
    const int k = k0 + tid;
    bool is_first = false;
    unsigned h = 0;
    if (k < K) {
      const int c = sm.conn[k];
      h = tg_hash(c);
      while (sm.hkey[h] != c) h = (h + 1) & (kTgHash - 1);
      is_first = sm.hmin[h] == k;
    }
    int tot;
    const int r = n + block_excl_scan(is_first ? 1 : 0, sm.scan, &tot);
    if (is_first) {
      if (r < QOT_TG_MAX_NODES) {
        sm.hnode[h] = r;
        sm.first[r] = k;
      } else {
        sm.bad = 1;
      }
    }
    n += tot;
  }
  __syncthreads();
  if (sm.bad || n > QOT_TG_MAX_NODES) { bad = true; n = 0; K = 0; }
  for (int k = tid; k < K; k += kTgThreads) {
    const int c = sm.conn[k];
    unsigned h = tg_hash(c);
    while (sm.hkey[h] != c) h = (h + 1) & (kTgHash - 1);
    sm.node[k] = static_cast<unsigned short>(sm.hnode[h]);
    if (k == 0 || sm.link[k - 1] != sm.link[k]) sm.seg[sm.link[k]] = k;       // start of the link's run
  }
  __syncthreads();

  K_out = K;
  n_out = n;
  bad_out = bad;
}

// per-sample record of the scan pass (qot_lightpath_graph_count): everything the pack pass needs, so the
// sample tensor is read ONCE
struct LpRecord {
  int n, E, pad0, pad1;
  int conn[QOT_TG_MAX_NODES];
  float x[QOT_TG_MAX_NODES][5];
  unsigned adj[QOT_TG_MAX_NODES][kTgAdjWords];
};

__global__ void __launch_bounds__(kTgThreads)
lp_graph_scan_kernel(const float* __restrict__ data, const double* __restrict__ freqs,
                     const double* __restrict__ target, qot_lp_graph_cfg_t cfg, int32_t* __restrict__ counts,
                     LpRecord* __restrict__ records, float* __restrict__ y, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) char tg_smem_raw[];
  TgSmem& sm = *reinterpret_cast<TgSmem*>(tg_smem_raw);
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int L = cfg.L, Q = cfg.Q, LQ = L * Q;
  const float* __restrict__ d = data + s * static_cast<int64_t>(cfg.F) * LQ;
  int K, n;
  bool bad;
  tg_scan_and_rank(sm, d, cfg, K, n, bad);
  (void)L;

  // ---- 4. interactions per link: entries of a link are contiguous and ordered by frequency index
  for (int i = tid; i < K; i += kTgThreads) {
    const int l = sm.link[i];
    const int a = sm.seg[l];
    int b = a;
    while (b < K && sm.link[b] == l) ++b;                                   // runs are short (<= Q)
    // links used by a single lightpath are skipped (to_graph.py:285-286)
    bool two = false;
    for (int j = a; j < b; ++j) two |= sm.node[j] != sm.node[a];
    if (!two) continue;
    const double fi = freqs[sm.freq[i]];
    const int ni = sm.node[i];
    for (int j = a; j < b; ++j) {
      const double df = fabs(fi - freqs[sm.freq[j]]);                        // :296, float64 like numpy
      if (df < cfg.freq_threshold && df > 0.0) atomicOr(&sm.adj[ni][sm.node[j] >> 5], 1u << (sm.node[j] & 31));
    }
  }
  __syncthreads();

  // ---- 5. directed edges, sorted by (source, target): one adjacency row per thread
  int deg = 0;
  if (tid < n)
    for (int w = 0; w < kTgAdjWords; ++w) deg += __popc(sm.adj[tid][w]);
  int E;
  const int e_off = block_excl_scan(deg, sm.scan, &E);
  (void)e_off;
  LpRecord& rec = records[s];
  if (tid == 0) {
    counts[2 * s] = n;
    counts[2 * s + 1] = E;
    rec.n = n;
    rec.E = E;
    if (bad) atomicOr(status, 1);
  }
  if (tid < n) {
#pragma unroll
    for (int w = 0; w < kTgAdjWords; ++w) rec.adj[tid][w] = sm.adj[tid][w];
    // ---- 3. node row from the node's first channel (dataset.py:74-80; sorted-name column order)
    const int k = sm.first[tid];
    const int ch = sm.link[k] * Q + sm.freq[k];
    const float osnr = d[static_cast<int64_t>(cfg.i_osnr) * LQ + ch], snr = d[static_cast<int64_t>(cfg.i_snr) * LQ + ch],
                ber = d[static_cast<int64_t>(cfg.i_ber) * LQ + ch];
    float* xr = rec.x[tid];
    constexpr int col[4] = {0, 2, 3, 4};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double v = static_cast<double>(d[static_cast<int64_t>(cfg.i_feat[q]) * LQ + ch]);
      xr[col[q]] = static_cast<float>((v - cfg.feat_lo[q]) / (cfg.feat_hi[q] - cfg.feat_lo[q]));
    }
    xr[1] = (osnr == -1.f && snr == -1.f && ber == -1.f) ? 1.0f : 0.0f;   // to_graph.py:247-251
    rec.conn[tid] = sm.conn[k];
  }
  if (tid < 3) {
    const double v = target[s * cfg.T + cfg.i_tgt[tid]];
    y[s * 3 + tid] = static_cast<float>((v - cfg.tgt_lo[tid]) / (cfg.tgt_hi[tid] - cfg.tgt_lo[tid]));   // dataset.py:111-121
  }
}

// pack pass: per-sample records -> the packed store (directed edges sorted by (source, target): one
// adjacency row per thread)
__global__ void __launch_bounds__(kTgThreads)
lp_graph_pack_kernel(const LpRecord* __restrict__ records, const int64_t* __restrict__ node_ptr,
                     const int64_t* __restrict__ edge_ptr, float* __restrict__ node_feat,
                     int64_t* __restrict__ conn_ids, int32_t* __restrict__ edge_src,
                     int32_t* __restrict__ edge_dst, int32_t* __restrict__ status) {
  __shared__ int warp_sums[kTgThreads / 32];
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const LpRecord& rec = records[s];
  const int n = rec.n;
  const int64_t n0 = node_ptr[s], e0 = edge_ptr[s];
  unsigned row[kTgAdjWords];
  int deg = 0;
#pragma unroll
  for (int w = 0; w < kTgAdjWords; ++w) {
    row[w] = tid < n ? rec.adj[tid][w] : 0u;
    deg += __popc(row[w]);
  }
  int E;
  const int e_off = block_excl_scan(deg, warp_sums, &E);
  if (node_ptr[s + 1] - n0 != n || edge_ptr[s + 1] - e0 != E || E != rec.E) {   // offsets from another pass / another tensor
    if (tid == 0) atomicOr(status, 2);
    return;
  }
  if (tid < n) {
    int64_t o = e0 + e_off;
#pragma unroll
    for (int w = 0; w < kTgAdjWords; ++w) {
      unsigned m = row[w];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        edge_src[o] = tid;
        edge_dst[o] = 32 * w + bit;
        ++o;
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) node_feat[(n0 + tid) * 5 + q] = rec.x[tid][q];
    conn_ids[n0 + tid] = rec.conn[tid];
  }
}


// Topological representation: to_graph.py::create_topological_graph (:62-184) + TopologicalDataset.__getitem__
// (topological_training/dataset.py:46-123).  Lightpaths are added to an nx.Graph in ascending conn_id
// (np.unique); one edge per node pair survives -- its adjacency position dates from the FIRST lightpath
// of the pair, its attributes from the LAST; dataset.py:57 relabels through a copy that re-orders every
// adjacency list (smaller neighbours first, ascending; then the rest in first-seen order); from_networkx
// lists each node's neighbours in that order.  All of it is ranking over <= 256 lightpaths: counting
// sorts in shared memory, bit-exact edge ORDER.
struct TpSmem {
  TgSmem base;
  int tord[QOT_TG_MAX_NODES];                 // time (ascending conn_id rank) -> node rank r
  short ea[QOT_TG_MAX_NODES], eb[QOT_TG_MAX_NODES];   // endpoints (0-based) of the lightpath added at time t
  short tfirst[QOT_TG_MAX_NODES], tlast[QOT_TG_MAX_NODES];
  int dkey[2 * QOT_TG_MAX_NODES];             // sort key of every directed edge candidate (-1: none)
};

struct TpRecord {
  int E, pad0, pad1, pad2;
  short src[2 * QOT_TG_MAX_NODES], dst[2 * QOT_TG_MAX_NODES];
  float feat[2 * QOT_TG_MAX_NODES][4];
};

__global__ void __launch_bounds__(kTgThreads)
tp_graph_scan_kernel(const float* __restrict__ data, const double* __restrict__ target, qot_lp_graph_cfg_t cfg,
                     int num_nodes, int i_src, int i_dst, int32_t* __restrict__ counts,
                     TpRecord* __restrict__ records, float* __restrict__ y, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) char tg_smem_raw[];
  TpSmem& tp = *reinterpret_cast<TpSmem*>(tg_smem_raw);
  TgSmem& sm = tp.base;
  const int tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int Q = cfg.Q, LQ = cfg.L * Q;
  const float* __restrict__ d = data + s * static_cast<int64_t>(cfg.F) * LQ;
  int K, n;
  bool bad;
  tg_scan_and_rank(sm, d, cfg, K, n, bad);
  (void)K;
  // time order = ascending conn_id (np.unique, to_graph.py:156): counting sort over the n distinct ids
  if (tid < n) {
    const int c = sm.conn[sm.first[tid]];
    int t = 0;
    for (int r = 0; r < n; ++r) t += sm.conn[sm.first[r]] < c;
    tp.tord[t] = tid;
  }
  __syncthreads();
  if (tid < n) {
    const int k = sm.first[tp.tord[tid]];
    const int ch = sm.link[k] * Q + sm.freq[k];
    const int a = static_cast<int>(d[static_cast<int64_t>(i_src) * LQ + ch]) - 1;     // nodes 1..75 -> 0..74
    const int b = static_cast<int>(d[static_cast<int64_t>(i_dst) * LQ + ch]) - 1;
    if (a < 0 || b < 0 || a >= num_nodes || b >= num_nodes) sm.bad = 1;
    tp.ea[tid] = static_cast<short>(a);
    tp.eb[tid] = static_cast<short>(b);
  }
  __syncthreads();
  if (sm.bad) { bad = true; n = 0; }
  // first / last lightpath of every unordered node pair
  if (tid < n) {
    const int lo = min(tp.ea[tid], tp.eb[tid]), hi = max(tp.ea[tid], tp.eb[tid]);
    int tf = tid, tl = tid;
    for (int t = 0; t < n; ++t) {
      if (min(tp.ea[t], tp.eb[t]) == lo && max(tp.ea[t], tp.eb[t]) == hi) {
        tf = min(tf, t);
        tl = max(tl, t);
      }
    }
    tp.tfirst[tid] = static_cast<short>(tf);
    tp.tlast[tid] = static_cast<short>(tl);
  }
  __syncthreads();
  // directed candidates: slot 2t = (a -> b), slot 2t+1 = (b -> a) of the pair's first lightpath.
  // key = (source, smaller-neighbour-first, then first-seen time): see the header comment
  for (int q = tid; q < 2 * QOT_TG_MAX_NODES; q += kTgThreads) {
    const int t = q >> 1;
    int key = -1;
    if (t < n && tp.tfirst[t] == t) {
      const int a = tp.ea[t], b = tp.eb[t];
      const int w = (q & 1) ? b : a, x = (q & 1) ? a : b;
      if (!((q & 1) && a == b)) key = (w << 12) | (x < w ? x : (1 << 11) | t);    // a self loop is listed once
    }
    tp.dkey[q] = key;
  }
  __syncthreads();
  int mine[2], pos[2] = {0, 0};
#pragma unroll
  for (int u = 0; u < 2; ++u) mine[u] = tp.dkey[tid + u * kTgThreads];
  int E = 0;
  for (int q = 0; q < 2 * n; ++q) {
    const int kq = tp.dkey[q];
    if (kq < 0) continue;
    ++E;
#pragma unroll
    for (int u = 0; u < 2; ++u) pos[u] += kq < mine[u];
  }
  TpRecord& rec = records[s];
  if (tid == 0) {
    counts[s] = E;
    rec.E = E;
    if (bad) atomicOr(status, 1);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (mine[u] < 0) continue;
    const int q = tid + u * kTgThreads, t = q >> 1;
    const int a = tp.ea[t], b = tp.eb[t];
    const int o = pos[u];
    rec.src[o] = static_cast<short>((q & 1) ? b : a);
    rec.dst[o] = static_cast<short>((q & 1) ? a : b);
    // attributes of the LAST lightpath of the pair, sorted-name order [freq, mod_order, num_spans, path_len]
    const int k = sm.first[tp.tord[tp.tlast[t]]];
    const int ch = sm.link[k] * Q + sm.freq[k];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const double v = static_cast<double>(d[static_cast<int64_t>(cfg.i_feat[f]) * LQ + ch]);
      rec.feat[o][f] = static_cast<float>((v - cfg.feat_lo[f]) / (cfg.feat_hi[f] - cfg.feat_lo[f]));
    }
  }
  if (tid < 3) {
    const double v = target[s * cfg.T + cfg.i_tgt[tid]];
    y[s * 3 + tid] = static_cast<float>((v - cfg.tgt_lo[tid]) / (cfg.tgt_hi[tid] - cfg.tgt_lo[tid]));
  }
}

__global__ void __launch_bounds__(kTgThreads)
tp_graph_pack_kernel(const TpRecord* __restrict__ records, const int64_t* __restrict__ edge_ptr,
                     int32_t* __restrict__ edge_src, int32_t* __restrict__ edge_dst,
                     float* __restrict__ edge_feat, int32_t* __restrict__ status) {
  const int64_t s = blockIdx.x;
  const TpRecord& rec = records[s];
  const int64_t e0 = edge_ptr[s];
  const int E = rec.E;
  if (edge_ptr[s + 1] - e0 != E) {
    if (threadIdx.x == 0) atomicOr(status, 2);
    return;
  }
  for (int o = threadIdx.x; o < E; o += kTgThreads) {
    edge_src[e0 + o] = rec.src[o];
    edge_dst[e0 + o] = rec.dst[o];
    reinterpret_cast<float4*>(edge_feat)[e0 + o] = *reinterpret_cast<const float4*>(rec.feat[o]);
  }
}

static int tg_check(const float* data, const double* freqs, int64_t S, const qot_lp_graph_cfg_t* cfg, const char* who) {
  QOT_REQUIRE(cfg && S >= 0, "%s: bad argument", who);
  QOT_REQUIRE(S == 0 || (data && freqs), "%s: null input", who);
  QOT_REQUIRE(cfg->F > 0 && cfg->L > 0 && cfg->Q > 0 && cfg->Q <= 65535 && cfg->L <= 65535 &&
                  static_cast<int64_t>(cfg->L) * cfg->Q < (1ll << 30), "%s: bad tensor extents", who);
  const int rows[8] = {cfg->i_conn, cfg->i_osnr, cfg->i_snr, cfg->i_ber, cfg->i_feat[0], cfg->i_feat[1], cfg->i_feat[2], cfg->i_feat[3]};
  for (int r : rows) QOT_REQUIRE(r >= 0 && r < cfg->F, "%s: lp_feat row index out of range", who);
  return QOT_OK;
}
static int tg_attr() {
  static std::atomic<unsigned long long> done{0};
  return once_per_device(done, [] {
    QOT_CUDA(cudaFuncSetAttribute(lp_graph_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TgSmem))));
    QOT_CUDA(cudaFuncSetAttribute(tp_graph_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TpSmem))));
    return static_cast<int>(QOT_OK);
  });
}

}  // namespace qot

using namespace qot;

extern "C" size_t qot_lightpath_graph_scratch_bytes(int64_t S) { return S > 0 ? static_cast<size_t>(S) * sizeof(LpRecord) : 256; }
extern "C" size_t qot_topological_graph_scratch_bytes(int64_t S) { return S > 0 ? static_cast<size_t>(S) * sizeof(TpRecord) : 256; }

extern "C" int qot_lightpath_graph_count(const float* data, const double* freqs, const double* target, int64_t S,
                                         const qot_lp_graph_cfg_t* cfg, int32_t* counts, float* y, void* scratch,
                                         size_t scratch_bytes, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = tg_check(data, freqs, S, cfg, "qot_lightpath_graph_count");
  if (rc) return rc;
  QOT_REQUIRE(status && (S == 0 || (counts && target && y)), "qot_lightpath_graph_count: null argument");
  QOT_REQUIRE(cfg->T > 0, "qot_lightpath_graph_count: bad target extent");
  for (int k = 0; k < 3; ++k) QOT_REQUIRE(cfg->i_tgt[k] >= 0 && cfg->i_tgt[k] < cfg->T, "qot_lightpath_graph_count: target column out of range");
  QOT_REQUIRE(scratch && scratch_bytes >= qot_lightpath_graph_scratch_bytes(S) && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
              "qot_lightpath_graph_count: scratch missing, misaligned or smaller than qot_lightpath_graph_scratch_bytes(S)");
  if (S == 0) return QOT_OK;
  if ((rc = tg_attr())) return rc;
  lp_graph_scan_kernel<<<static_cast<unsigned>(S), kTgThreads, sizeof(TgSmem), stream>>>(
      data, freqs, target, *cfg, counts, static_cast<LpRecord*>(scratch), y, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_lightpath_graph_fill(const void* scratch, int64_t S, const int64_t* node_ptr,
                                        const int64_t* edge_ptr, float* node_feat, int64_t* conn_ids,
                                        int32_t* edge_src, int32_t* edge_dst, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(S >= 0 && status && (S == 0 || (scratch && node_ptr && edge_ptr && node_feat && conn_ids && edge_src && edge_dst)),
              "qot_lightpath_graph_fill: null argument");
  if (S == 0) return QOT_OK;
  lp_graph_pack_kernel<<<static_cast<unsigned>(S), kTgThreads, 0, stream>>>(
      static_cast<const LpRecord*>(scratch), node_ptr, edge_ptr, node_feat, conn_ids, edge_src, edge_dst, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_topological_graph_count(const float* data, const double* target, int64_t S,
                                           const qot_lp_graph_cfg_t* cfg, int32_t num_nodes, int32_t i_src,
                                           int32_t i_dst, int32_t* counts, float* y, void* scratch,
                                           size_t scratch_bytes, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = tg_check(data, reinterpret_cast<const double*>(data), S, cfg, "qot_topological_graph_count");
  if (rc) return rc;
  QOT_REQUIRE(status && (S == 0 || (counts && target && y)), "qot_topological_graph_count: null argument");
  QOT_REQUIRE(num_nodes > 0 && num_nodes < 2048 && i_src >= 0 && i_src < cfg->F && i_dst >= 0 && i_dst < cfg->F,
              "qot_topological_graph_count: bad node count or endpoint rows");
  QOT_REQUIRE(cfg->T > 0, "qot_topological_graph_count: bad target extent");
  for (int k = 0; k < 3; ++k) QOT_REQUIRE(cfg->i_tgt[k] >= 0 && cfg->i_tgt[k] < cfg->T, "qot_topological_graph_count: target column out of range");
  QOT_REQUIRE(scratch && scratch_bytes >= qot_topological_graph_scratch_bytes(S) && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
              "qot_topological_graph_count: scratch missing, misaligned or smaller than qot_topological_graph_scratch_bytes(S)");
  if (S == 0) return QOT_OK;
  if ((rc = tg_attr())) return rc;
  tp_graph_scan_kernel<<<static_cast<unsigned>(S), kTgThreads, sizeof(TpSmem), stream>>>(
      data, target, *cfg, num_nodes, i_src, i_dst, counts, static_cast<TpRecord*>(scratch), y, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_topological_graph_fill(const void* scratch, int64_t S, const int64_t* edge_ptr,
                                          int32_t* edge_src, int32_t* edge_dst, float* edge_feat,
                                          int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(S >= 0 && status && (S == 0 || (scratch && edge_ptr && edge_src && edge_dst && edge_feat)),
              "qot_topological_graph_fill: null argument");
  QOT_REQUIRE((reinterpret_cast<uintptr_t>(edge_feat) & 15) == 0, "qot_topological_graph_fill: edge_feat must be 16-byte aligned");
  if (S == 0) return QOT_OK;
  tp_graph_pack_kernel<<<static_cast<unsigned>(S), kTgThreads, 0, stream>>>(
      static_cast<const TpRecord*>(scratch), edge_ptr, edge_src, edge_dst, edge_feat, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
