// Integer side of the hot path: device-wide scan, stable CSR construction,
// graph offsets and the device-side collate.  Everything here is bit-exact
// against oracle.build_csr_ref / oracle.collate_ref.
//
// Replaces: PyG Batch.from_data_list driven by the reference DataLoaders
// (topological_training/train.py:93-95, lightpath_training/train.py:94-96) and the
// scatter indexing inside MessagePassing.propagate.
#include <algorithm>

#include "common.cuh"

namespace qot {

// ===========================================================================
// Single-pass exclusive scan (decoupled look-back).  One 64-bit word per tile
// carries {flag, value} so no fence is needed between them.
// ===========================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

size_t scan_workspace_bytes(int64_t n) {
  int64_t tiles = cdiv(n + 1, kScanTile);
  return align_up(256 + static_cast<size_t>(tiles) * sizeof(unsigned long long));
}

__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const int32_t* __restrict__ in, int32_t add, int32_t* __restrict__ out, int64_t n,
            unsigned int* ticket, volatile unsigned long long* state) {
  __shared__ unsigned int s_tile;
  __shared__ int32_t s_warp[kScanThreads / 32];
  __shared__ int32_t s_prefix;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const int64_t base = static_cast<int64_t>(tile) * kScanTile + static_cast<int64_t>(tid) * kScanItems;

  int32_t v[kScanItems];
  int32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = (idx < n) ? (in[idx] + add) : 0;
    tsum += v[i];
  }
  // block exclusive scan of the per-thread sums
  int32_t incl = tsum;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int32_t warp_off = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    int32_t t = s_warp[w];
    if (w < warp) warp_off += t;
    block_total += t;
  }
  const int32_t thread_excl = warp_off + incl - tsum;

  if (tid == 0) {
    int32_t prefix = 0;
    if (tile == 0) {
      state[0] = (2ull << 32) | static_cast<unsigned int>(block_total);
    } else {
      state[tile] = (1ull << 32) | static_cast<unsigned int>(block_total);
      int64_t t = static_cast<int64_t>(tile) - 1;
      while (true) {
        unsigned long long s = state[t];
        unsigned int flag = static_cast<unsigned int>(s >> 32);
        if (flag == 0) continue;   // predecessor (older ticket) not published yet
        prefix += static_cast<int32_t>(static_cast<unsigned int>(s));
        if (flag == 2) break;
        --t;
      }
      state[tile] = (2ull << 32) | static_cast<unsigned int>(prefix + block_total);
    }
    s_prefix = prefix;
  }
  __syncthreads();
  int32_t run = s_prefix + thread_excl;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx <= n) out[idx] = run;
    run += v[i];
  }
}

int exclusive_scan_i32(const int32_t* in, int32_t add, int32_t* out, int64_t n, void* ws,
                       cudaStream_t stream) {
  const int64_t tiles = cdiv(n + 1, kScanTile);
  QOT_CUDA(cudaMemsetAsync(ws, 0, scan_workspace_bytes(n), stream));
  unsigned int* ticket = static_cast<unsigned int*>(ws);
  auto* state = reinterpret_cast<volatile unsigned long long*>(static_cast<char*>(ws) + 256);
  scan_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, stream>>>(in, add, out, n, ticket, state);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

// ===========================================================================
// Stable CSR: histogram -> scan -> unordered fill -> per-row sort by edge id.
// Integer atomics only decide *where* inside a row an edge lands first; the row
// sort makes the final layout independent of that, so the result is
// deterministic and equals a stable argsort by the key row.
// ===========================================================================
constexpr int kSmallRow = 64;   // rows up to this in-degree are insertion-sorted by one thread

__global__ void csr_count_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ other,
                                 int64_t E, int64_t N, int drop_self, int32_t* __restrict__ deg,
                                 int32_t* status) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < E;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t d = key[e], o = other[e];
    if (static_cast<uint64_t>(d) >= static_cast<uint64_t>(N) ||
        static_cast<uint64_t>(o) >= static_cast<uint64_t>(N)) {
      if (status) *status = 1;
      continue;
    }
    if (drop_self && d == o) continue;
    atomicAdd(&deg[d], 1);
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ other,
                                int64_t E, int64_t N, int drop_self,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ deg,
                                int32_t* __restrict__ eid) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < E;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t d = key[e], o = other[e];
    if (static_cast<uint64_t>(d) >= static_cast<uint64_t>(N) ||
        static_cast<uint64_t>(o) >= static_cast<uint64_t>(N))
      continue;
    if (drop_self && d == o) continue;
    const int32_t slot = atomicSub(&deg[d], 1) - 1;   // deg returns to 0: reusable as scratch
    eid[rowptr[d] + slot] = static_cast<int32_t>(e);
  }
}

__global__ void csr_sort_rows_kernel(const int64_t* __restrict__ other, int64_t E, int64_t N,
                                     int add_self, const int32_t* __restrict__ rowptr,
                                     int32_t* __restrict__ eid, int32_t* __restrict__ nbr,
                                     int32_t* __restrict__ big_count, int32_t* __restrict__ big_rows) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= N) return;
  const int32_t beg = rowptr[r];
  const int32_t end = rowptr[r + 1] - (add_self ? 1 : 0);
  const int32_t n = end - beg;
  if (add_self) {
    eid[end] = static_cast<int32_t>(E + r);
    nbr[end] = static_cast<int32_t>(r);
  }
  if (n > kSmallRow) {
    big_rows[atomicAdd(big_count, 1)] = static_cast<int32_t>(r);
    return;
  }
  for (int32_t i = beg + 1; i < end; ++i) {
    const int32_t v = eid[i];
    int32_t j = i - 1;
    while (j >= beg && eid[j] > v) {
      eid[j + 1] = eid[j];
      --j;
    }
    eid[j + 1] = v;
  }
  for (int32_t i = beg; i < end; ++i) nbr[i] = static_cast<int32_t>(other[eid[i]]);
}

// Hub rows (more than kSmallRow entries): up to kBitonicMax entries are sorted by a bitonic network in
// shared memory; longer rows by rank (edge ids are unique), using the not-yet-written nbr segment of
// the row as the temporary copy.
constexpr int kBitonicMax = 4096;

__global__ void __launch_bounds__(256)
csr_sort_big_rows_kernel(const int64_t* __restrict__ other, int add_self,
                         const int32_t* __restrict__ rowptr, int32_t* __restrict__ eid,
                         int32_t* __restrict__ nbr, const int32_t* __restrict__ big_count,
                         const int32_t* __restrict__ big_rows) {
  __shared__ int32_t s_key[kBitonicMax];
  const int32_t nbig = *big_count;
  for (int32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
    const int32_t r = big_rows[b];
    const int32_t beg = rowptr[r];
    const int32_t end = rowptr[r + 1] - (add_self ? 1 : 0);
    const int32_t n = end - beg;
    if (n <= kBitonicMax) {
      int32_t m = 1;
      while (m < n) m <<= 1;
      for (int32_t i = threadIdx.x; i < m; i += blockDim.x) s_key[i] = i < n ? eid[beg + i] : 0x7fffffff;
      __syncthreads();
      for (int32_t k = 2; k <= m; k <<= 1) {
        for (int32_t j = k >> 1; j > 0; j >>= 1) {
          for (int32_t i = threadIdx.x; i < m; i += blockDim.x) {
            const int32_t p = i ^ j;
            if (p > i) {
              const int32_t a = s_key[i], c = s_key[p];
              const bool up = (i & k) == 0;
              if ((a > c) == up) {
                s_key[i] = c;
                s_key[p] = a;
              }
            }
          }
          __syncthreads();
        }
      }
      for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int32_t e = s_key[i];
        eid[beg + i] = e;
        nbr[beg + i] = static_cast<int32_t>(other[e]);
      }
      __syncthreads();
      continue;
    }
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) nbr[beg + i] = eid[beg + i];
    __syncthreads();
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int32_t v = nbr[beg + i];
      int32_t rank = 0;
      for (int32_t j = 0; j < n; ++j) rank += (nbr[beg + j] < v) ? 1 : 0;
      eid[beg + rank] = v;
    }
    __syncthreads();
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x)
      nbr[beg + i] = static_cast<int32_t>(other[eid[beg + i]]);
    __syncthreads();
  }
}

static size_t big_list_entries(int64_t N, int64_t E) { return static_cast<size_t>((E + N) / kSmallRow + 2); }

// ptr[g] = first position whose graph id is >= g, for a sorted id sequence.  With
// `via` the id of position n is batch[via[n]] (edges -> graph of their source node)
// and a non-monotonic sequence raises *status.
__global__ void graph_ptr_kernel(const int64_t* __restrict__ batch, const int64_t* __restrict__ via,
                                 int64_t n_items, int64_t n_nodes, int64_t B,
                                 int64_t* __restrict__ ptr, int32_t* status) {
  const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (n > n_items) return;
  auto id_at = [&](int64_t i) -> int64_t {
    if (!via) return batch[i];
    const int64_t node = via[i];
    const int64_t peer = via[n_items + i];          // destination row of edge_index [2, E]
    if (static_cast<uint64_t>(node) >= static_cast<uint64_t>(n_nodes) ||
        static_cast<uint64_t>(peer) >= static_cast<uint64_t>(n_nodes)) {
      if (status) *status = 1;                      // an endpoint outside [0, N): not a valid batch
      return 0;
    }
    if (batch[peer] != batch[node] && status) *status = 1;   // an edge must stay inside one graph
    return batch[node];
  };
  const int64_t prev = (n == 0) ? -1 : id_at(n - 1);
  const int64_t cur = (n == n_items) ? B : id_at(n);
  if (cur < prev && status) *status = 1;
  const int64_t hi = cur < B ? cur : B;
  for (int64_t g = prev + 1; g <= hi; ++g) ptr[g] = n;
}

// One warp per graph: flat, coalesced copies of its node / edge slabs.
__global__ void __launch_bounds__(256)
collate_kernel(qot_store_t st, const int64_t* __restrict__ graph_ids, int64_t B,
               const int64_t* __restrict__ out_ptr, const int64_t* __restrict__ out_eptr,
               int64_t E, float* __restrict__ x, int64_t* __restrict__ edge_index,
               float* __restrict__ edge_attr, int64_t* __restrict__ batch,
               int64_t* __restrict__ node_ids, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (b >= B) return;
  const int64_t g = graph_ids[b];
  const int64_t n0 = st.node_ptr[g], n = st.node_ptr[g + 1] - n0;
  const int64_t e0 = st.edge_ptr[g], ne = st.edge_ptr[g + 1] - e0;
  const int64_t on = out_ptr[b], oe = out_eptr[b];
  if (x && st.node_feat) {
    const float* s = st.node_feat + n0 * st.node_dim;
    float* d = x + on * st.node_dim;
    for (int64_t i = lane; i < n * st.node_dim; i += 32) d[i] = s[i];
  }
  for (int64_t i = lane; i < n; i += 32) {
    if (batch) batch[on + i] = b;
    if (node_ids) node_ids[on + i] = i;
  }
  if (edge_index) {
    for (int64_t i = lane; i < ne; i += 32) {
      edge_index[oe + i] = st.edge_src[e0 + i] + on;
      edge_index[E + oe + i] = st.edge_dst[e0 + i] + on;
    }
  }
  if (edge_attr && st.edge_feat) {
    const float* s = st.edge_feat + e0 * st.edge_dim;
    float* d = edge_attr + oe * st.edge_dim;
    for (int64_t i = lane; i < ne * st.edge_dim; i += 32) d[i] = s[i];
  }
  if (y && st.y) {
    for (int i = lane; i < st.y_dim; i += 32) y[b * st.y_dim + i] = st.y[g * st.y_dim + i];
  }
}

}  // namespace qot

using namespace qot;

extern "C" size_t qot_csr_workspace_bytes(int64_t N, int64_t E) {
  if (N < 0 || E < 0) return 0;
  return align_up(static_cast<size_t>(N + 1) * 4) + scan_workspace_bytes(N) + align_up(256) +
         align_up(big_list_entries(N, E) * 4) + 256;
}

extern "C" int qot_build_csr(const int64_t* edge_index, int64_t E, int64_t N, int by, int flags,
                             int32_t* rowptr, int32_t* nbr, int32_t* eid, int32_t* status,
                             void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && E >= 0, "qot_build_csr: negative size");
  QOT_REQUIRE(E + N < (1ll << 31) - 1, "qot_build_csr: E+N exceeds int32 CSR range");
  QOT_REQUIRE(by == 0 || by == 1, "qot_build_csr: by must be 0 (source) or 1 (destination)");
  const int64_t max_entries = E + ((flags & 2) ? N : 0);
  QOT_REQUIRE(rowptr && (max_entries == 0 || (nbr && eid)), "qot_build_csr: null output");
  QOT_REQUIRE(E == 0 || edge_index, "qot_build_csr: null edge_index");
  QOT_REQUIRE(ws && ws_bytes >= qot_csr_workspace_bytes(N, E), "qot_build_csr: workspace too small");
  const int drop_self = flags & 1, add_self = (flags >> 1) & 1;
  const int64_t* key = edge_index + (by ? E : 0);
  const int64_t* other = edge_index + (by ? 0 : E);

  Carver c(ws);
  int32_t* deg = c.take<int32_t>(N + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(N));
  int32_t* big_count = c.take<int32_t>(64);
  int32_t* big_rows = c.take<int32_t>(big_list_entries(N, E));

  QOT_CUDA(cudaMemsetAsync(deg, 0, static_cast<size_t>(N + 1) * 4, stream));
  QOT_CUDA(cudaMemsetAsync(big_count, 0, 256, stream));
  if (status) QOT_CUDA(cudaMemsetAsync(status, 0, 4, stream));
  const int threads = 256;
  if (E > 0) {
    const int64_t blocks = std::min<int64_t>(cdiv(E, threads), kNumSMs * 16);
    csr_count_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(key, other, E, N, drop_self, deg, status);
    QOT_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i32(deg, add_self, rowptr, N, scan_ws, stream);
  if (rc) return rc;
  if (E > 0) {
    const int64_t blocks = std::min<int64_t>(cdiv(E, threads), kNumSMs * 16);
    csr_fill_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(key, other, E, N, drop_self, rowptr, deg, eid);
    QOT_LAUNCH_CHECK();
  }
  if (N > 0) {
    csr_sort_rows_kernel<<<static_cast<unsigned>(cdiv(N, threads)), threads, 0, stream>>>(
        other, E, N, add_self, rowptr, eid, nbr, big_count, big_rows);
    QOT_LAUNCH_CHECK();
    csr_sort_big_rows_kernel<<<32, 256, 0, stream>>>(other, add_self, rowptr, eid, nbr, big_count, big_rows);
    QOT_LAUNCH_CHECK();
  }
  return QOT_OK;
}

extern "C" int qot_graph_ptr(const int64_t* batch, int64_t N, int64_t B, int64_t* gptr, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && B >= 0 && gptr, "qot_graph_ptr: bad argument");
  QOT_REQUIRE(N == 0 || batch, "qot_graph_ptr: null batch");
  graph_ptr_kernel<<<static_cast<unsigned>(cdiv(N + 1, 256)), 256, 0, stream>>>(batch, nullptr, N, N, B, gptr, nullptr);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_edge_ptr(const int64_t* edge_index, int64_t E, const int64_t* batch, int64_t N,
                            int64_t B, int64_t* eptr, int32_t* status, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(N >= 0 && B >= 0 && E >= 0 && eptr, "qot_edge_ptr: bad argument");
  QOT_REQUIRE(E == 0 || (edge_index && batch), "qot_edge_ptr: null input");
  if (status) QOT_CUDA(cudaMemsetAsync(status, 0, 4, stream));
  graph_ptr_kernel<<<static_cast<unsigned>(cdiv(E + 1, 256)), 256, 0, stream>>>(batch, edge_index, E, N, B, eptr, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" int qot_collate(const qot_store_t* store, const int64_t* graph_ids, int64_t B,
                           const int64_t* out_ptr, const int64_t* out_eptr, int64_t N, int64_t E,
                           float* x, int64_t* edge_index, float* edge_attr, int64_t* batch,
                           int64_t* node_ids, float* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(store && store->node_ptr && store->edge_ptr, "qot_collate: null store");
  QOT_REQUIRE(B >= 0 && N >= 0 && E >= 0, "qot_collate: negative size");
  if (B == 0) return QOT_OK;
  QOT_REQUIRE(graph_ids && out_ptr && out_eptr, "qot_collate: null index arrays");
  QOT_REQUIRE(!edge_index || (store->edge_src && store->edge_dst), "qot_collate: store has no edges");
  const int threads = 256;
  const int64_t blocks = cdiv(B * 32, threads);
  collate_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      *store, graph_ids, B, out_ptr, out_eptr, E, x, edge_index, edge_attr, batch, node_ids, y);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
