// LightpathGNN eval forward, generations 0 and 1 (kept behind qot_lightpath_set_variant for comparison
// and as parity cross-checks of the current kernels in lightpath_infer.cu):
//   variant 0, lp_infer_kernel     : one warp per graph, weights staged per block with cp.async;
//   variant 1, lp_infer_sub_kernel : 8 lanes per graph, slabs staged with 4-byte cp.async, FFMA head.
#include <algorithm>

#include "lightpath_common.cuh"

namespace qot {

#define LP_TRACE(slot) do {} while (0)

__global__ void __launch_bounds__(kThreads, QOT_LP_OCC)
lp_infer_kernel(const float* __restrict__ x, const int64_t* __restrict__ esrc,
                const int64_t* __restrict__ edst,
                const int64_t* __restrict__ gptr, const int64_t* __restrict__ eptr,
                const int64_t* __restrict__ lptr, int64_t N, int64_t B,
                const float* __restrict__ prep, int lut_col, float* __restrict__ out,
                int64_t* __restrict__ lut_batch, int32_t* __restrict__ lut_node,
                int32_t* __restrict__ n_lut, int32_t* __restrict__ status) {
  __shared__ __align__(16) float s_w[kWeightFloats];            // projection + head weights (19.5 KB)
  __shared__ __align__(16) float s_x[kIW][kXF];
  __shared__ __align__(16) float s_y[kIW][kHC];
  __shared__ float s_z[kIW][32];
  __shared__ int s_msg[kIW][kMsgCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * kIW + warp;
  const bool active = g < B;

  LP_TRACE(0);
  // ---- (0) weights -> shared memory, asynchronously (lands while the graph data is in flight)
  for (int i = threadIdx.x * 4; i < kWeightFloats; i += kThreads * 4) cp_async16(s_w + i, prep + kOffWf + i);

  // ---- (1) graph extents: 6 lanes fetch gptr[g..g+1], eptr[g..g+1], lut_ptr[g..g+1]
  long long pv = 0;
  if (active && lane < 6) pv = (lane < 2) ? gptr[g + lane] : (lane < 4) ? eptr[g + lane - 2] : lptr[g + lane - 4];
  const int64_t n0 = __shfl_sync(kFull, pv, 0), n1 = __shfl_sync(kFull, pv, 1);
  const int64_t e0 = __shfl_sync(kFull, pv, 2), e1 = __shfl_sync(kFull, pv, 3);
  const int64_t l0 = __shfl_sync(kFull, pv, 4), l1 = __shfl_sync(kFull, pv, 5);
  const bool fast = active && n1 >= n0 && e1 >= e0 && (n1 - n0) <= kMaxN && (e1 - e0) <= 32 * kEC;
  const int n = fast ? static_cast<int>(n1 - n0) : 0;
  const int ne = fast ? static_cast<int>(e1 - e0) : 0;
  const int n0i = static_cast<int>(n0);          // N < 2^31 (checked on the host)
  if (g == B - 1 && lane == 0) n_lut[0] = static_cast<int32_t>(l1);

  LP_TRACE(1);
  // ---- (2) every compulsory byte of the graph requested at once
  unsigned dlp[kEC / 4] = {0xffffffffu, 0xffffffffu};   // graph-local destination per held edge, one byte
                                                        // each (0xff: none / outside the slab)
  if (fast) {
    const float* __restrict__ xg = x + n0 * kF + lane;
    const int64_t* __restrict__ dg = edst + e0 + lane;
    const int nf = n * kF;
    float xr[kXR];
    long long dv[kEC];
#pragma unroll
    for (int k = 0; k < kXR; ++k) xr[k] = (lane + 32 * k < nf) ? __ldg(xg + 32 * k) : 0.f;
#pragma unroll
    for (int k = 0; k < kEC; ++k) dv[k] = (lane + 32 * k < ne) ? dg[32 * k] : -1ll;
#pragma unroll
    for (int k = 0; k < kXR; ++k)
      if (32 * k < nf) s_x[warp][lane + 32 * k] = xr[k];
#pragma unroll
    for (int k = 0; k < kEC; ++k) {
      const unsigned lo = static_cast<unsigned>(dv[k]);
      const unsigned hi = static_cast<unsigned>(static_cast<unsigned long long>(dv[k]) >> 32);
      const unsigned loc = lo - static_cast<unsigned>(n0i);
      const unsigned byte = (hi == 0u && loc < static_cast<unsigned>(n)) ? loc : 0xffu;
      dlp[k >> 2] = (dlp[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (byte << (8 * (k & 3)));
    }
  }
  cp_async_wait_all();
  __syncthreads();                                               // s_w (all threads) and s_x are in place
  LP_TRACE(2);

  // ---- (3) LUT rows of this graph, in ascending node order, at the offset lut_ptr gives
  int64_t orow = l0;
  int found = 0;
  if (fast) {
    const float* sx = s_x[warp];
    int* msg = s_msg[warp];
    unsigned a0 = __ballot_sync(kFull, lane < n && sx[lane * kF + lut_col] == 1.0f);
    unsigned a1 = __ballot_sync(kFull, lane + 32 < n && sx[(lane + 32) * kF + lut_col] == 1.0f);
    found = __popc(a0) + __popc(a1);
    if (found == l1 - l0) {
      const int h = lane & 3;
      float As[kF], Ad[kF];
#pragma unroll
      for (int k = 0; k < kF; ++k) {
        As[k] = __ldg(prep + kOffAsrc + k * kHeads + h);
        Ad[k] = __ldg(prep + kOffAdst + k * kHeads + h);
      }
      while (a0 | a1) {
        int il;
        if (a0) { il = __ffs(a0) - 1; a0 &= a0 - 1; } else { il = 32 + __ffs(a1) - 1; a1 &= a1 - 1; }
        // pass 1: local ids of the edges that point at the LUT node, in edge order
        int mc = 0;
#pragma unroll
        for (int k = 0; k < kEC; ++k) {
          if (32 * k < ne) {                                     // warp-uniform
            const bool hit = ((dlp[k >> 2] >> (8 * (k & 3))) & 0xffu) == static_cast<unsigned>(il);
            const unsigned hm = __ballot_sync(kFull, hit);
            if (hit && mc < kMsgCap - 32) msg[mc + __popc(hm & ((1u << lane) - 1u))] = lane + 32 * k;
            mc += __popc(hm);
          }
        }
        __syncwarp();
        float ov;
        bool done = false;
        if (mc <= 32) {                                          // warp-uniform; hubs take the generic path
          // pass 2: their sources (one gather); self loops and out-of-range ids dropped, order kept
          long long sj = -1;
          if (lane < mc) sj = esrc[e0 + msg[lane]];
          const bool inN = static_cast<uint64_t>(sj) < static_cast<uint64_t>(N);
          const long long sloc = sj - n0;
          const bool inslab = sloc >= 0 && sloc < n;
          const bool ok = inslab && sloc != il;
          const unsigned outside = __ballot_sync(kFull, inN && !inslab);
          const unsigned hm = __ballot_sync(kFull, ok);
          __syncwarp();
          if (ok) msg[__popc(hm & ((1u << lane) - 1u))] = static_cast<int>(sloc);
          mc = __popc(hm);
          if (outside == 0) {
            if (lane == 0) msg[mc] = il;                         // the appended self loop comes last
            ++mc;
            __syncwarp();
            float d_i = 0.f;
#pragma unroll
            for (int k = 0; k < kF; ++k) d_i = fmaf(sx[il * kF + k], Ad[k], d_i);
            AttnState as;
            attn_consume(as, msg, mc, [&](int j, int k) { return sx[j * kF + k]; }, As, d_i, lane);
            attn_finish(as, s_z[warp], lane);
            __syncwarp();
            ov = lut_row_head(s_w, s_z[warp], s_y[warp], lane);
            __syncwarp();
            done = true;
          }
        }
        if (!done)
          ov = lut_row_global(x, esrc, edst, e0, e1, N, n0 + il, prep, s_w, msg, s_z[warp], s_y[warp], lane);
        if (lane < QOT_OUT) out[orow * QOT_OUT + lane] = ov;
        if (lane == 0) {
          lut_batch[orow] = g;
          lut_node[orow] = static_cast<int32_t>(n0 + il);
        }
        ++orow;
      }
    }
  } else if (active) {
    for (int64_t nb = n0; nb < n1; nb += 32) {
      const int64_t node = nb + lane;
      unsigned mask = __ballot_sync(kFull, node < n1 && x[node * kF + lut_col] == 1.0f);
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        ++found;
        if (orow < l1) {                                         // never write past this graph's rows
          const int64_t i = nb + bit;
          const float ov = lut_row_global(x, esrc, edst, e0, e1, N, i, prep, s_w, s_msg[warp], s_z[warp], s_y[warp], lane);
          if (lane < QOT_OUT) out[orow * QOT_OUT + lane] = ov;
          if (lane == 0) {
            lut_batch[orow] = g;
            lut_node[orow] = static_cast<int32_t>(i);
          }
          ++orow;
        }
      }
    }
  }
  // lut_ptr must describe THIS x: a stale / foreign lut_ptr is reported, never trusted silently
  if (active && lane == 0 && found != l1 - l0) atomicOr(status, 1);
  LP_TRACE(7);
}

// =====================================================================================
// Sub-warp variant.  The per-graph scan (find the LUT node, find its in-edges, gather their
// sources) and the attention need far fewer than 32 lanes per ~32-node graph, and warp-wide
// instructions cost the same however few lanes do useful work.  Here a graph gets 8 lanes
// (4 graphs per warp, 32 graphs per 256-thread block) for that phase; afterwards the block
// runs the 128 -> 32 -> 3 readout heads on full warps, two rows at a time so every weight read
// from shared memory serves two rows.  Graphs outside the fast path (n > 64, E_g > 256, not
// exactly one LUT node, more than 15 in-edges on it, a source outside the slab) are handed to
// the generic per-warp path at the end of the same launch.
// =====================================================================================
#ifndef QOT_LP_SUB_OCC
#define QOT_LP_SUB_OCC 4
#endif
constexpr int kSlabNodes = 1216;          // nodes of one block's 32 graphs staged together (mean 1024 at
                                          // n ~ U{8..56}; graphs that do not fit take the generic path)
struct SubSmem {
  float x[kSlabNodes * kF];               // the block's node slabs, packed as in global memory, 23.8 KB
  float z[kGPB][kHeads * kF];             // attention outputs, z[h*5+f]
  float y[8][2][kHC];                     // per warp: two activation rows (generic path scratch too)
  int msg[kGPB][kSubMsg];
  SubMeta meta[kGPB];
};

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}

// readout head for two rows at once (za / zb: z[h*5+f]); lanes 0..2 get out[k] of each row
__device__ __forceinline__ void lut_head2(const float* __restrict__ w, const float* za, const float* zb,
                                          float* ya, float* yb, int lane, float& oa, float& ob) {
  constexpr int oShift = kOffShift - kOffWf, oW1 = kOffW1t - kOffWf, oB1 = kOffB1 - kOffWf,
                oW2 = kOffW2 - kOffWf, oB2 = kOffB2 - kOffWf;
#pragma unroll
  for (int h2 = 0; h2 < kHeads; ++h2) {
    float va = w[oShift + h2 * kC + lane], vb = va;
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      const float wv = w[(h2 * kF + k) * kC + lane];
      va = fmaf(wv, za[h2 * kF + k], va);
      vb = fmaf(wv, zb[h2 * kF + k], vb);
    }
    ya[h2 * kC + lane] = fmaxf(va, 0.f);
    yb[h2 * kC + lane] = fmaxf(vb, 0.f);
  }
  __syncwarp();
  const int cg = lane >> 3, og = lane & 7;
  const float4* __restrict__ w1 = reinterpret_cast<const float4*>(w + oW1) + cg * 32 * (kHid / 4) + og;
  const float4* __restrict__ qa = reinterpret_cast<const float4*>(ya + cg * 32);
  const float4* __restrict__ qb = reinterpret_cast<const float4*>(yb + cg * 32);
  float ha[4] = {0.f, 0.f, 0.f, 0.f}, hb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 a4 = qa[c4], b4 = qb[c4];
    const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 wv = w1[(4 * c4 + t) * (kHid / 4)];
      ha[0] = fmaf(wv.x, av[t], ha[0]); ha[1] = fmaf(wv.y, av[t], ha[1]);
      ha[2] = fmaf(wv.z, av[t], ha[2]); ha[3] = fmaf(wv.w, av[t], ha[3]);
      hb[0] = fmaf(wv.x, bv[t], hb[0]); hb[1] = fmaf(wv.y, bv[t], hb[1]);
      hb[2] = fmaf(wv.z, bv[t], hb[2]); hb[3] = fmaf(wv.w, bv[t], hb[3]);
    }
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      ha[t] += __shfl_xor_sync(kFull, ha[t], o);
      hb[t] += __shfl_xor_sync(kFull, hb[t], o);
    }
  }
  const float4 b1 = reinterpret_cast<const float4*>(w + oB1)[og];
  const float b1v[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    ha[t] += b1v[t];
    hb[t] += b1v[t];
    ha[t] = ha[t] > 0.f ? ha[t] : 0.01f * ha[t];
    hb[t] = hb[t] > 0.f ? hb[t] : 0.01f * hb[t];
  }
  float pa[QOT_OUT], pb[QOT_OUT];
#pragma unroll
  for (int k = 0; k < QOT_OUT; ++k) {
    const float4 wk = reinterpret_cast<const float4*>(w + oW2 + k * kHid)[og];
    pa[k] = wk.x * ha[0] + wk.y * ha[1] + wk.z * ha[2] + wk.w * ha[3];
    pb[k] = wk.x * hb[0] + wk.y * hb[1] + wk.z * hb[2] + wk.w * hb[3];
  }
#pragma unroll
  for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
    for (int k = 0; k < QOT_OUT; ++k) {
      pa[k] += __shfl_xor_sync(kFull, pa[k], o);
      pb[k] += __shfl_xor_sync(kFull, pb[k], o);
    }
  }
  const float b2 = lane < QOT_OUT ? w[oB2 + lane] : 0.f;
  oa = (lane == 0 ? pa[0] : lane == 1 ? pa[1] : pa[2]) + b2;
  ob = (lane == 0 ? pb[0] : lane == 1 ? pb[1] : pb[2]) + b2;
}

__global__ void __launch_bounds__(256, QOT_LP_SUB_OCC)
lp_infer_sub_kernel(const float* __restrict__ x, const int64_t* __restrict__ esrc,
                    const int64_t* __restrict__ edst, const int64_t* __restrict__ gptr,
                    const int64_t* __restrict__ eptr, const int64_t* __restrict__ lptr, int64_t N,
                    int64_t B, const float* __restrict__ prep, int lut_col, float* __restrict__ out,
                    int64_t* __restrict__ lut_batch, int32_t* __restrict__ lut_node,
                    int32_t* __restrict__ n_lut, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) char sub_smem_raw[];
  SubSmem& sm = *reinterpret_cast<SubSmem*>(sub_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sg = lane >> 3, sl = lane & 7, base = lane & ~7;
  const int gl = warp * 4 + sg;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * kGPB + gl;
  const bool active = g < B;

  // ---- extents of the sub-group's graph (+ the block's first node: slabs are packed from there)
  const int64_t nb0 = gptr[min(static_cast<int64_t>(blockIdx.x) * kGPB, B)];
  long long pv = 0;
  if (active && sl < 6) pv = (sl < 2) ? gptr[g + sl] : (sl < 4) ? eptr[g + sl - 2] : lptr[g + sl - 4];
  const int64_t n0 = __shfl_sync(kFull, pv, base + 0), n1 = __shfl_sync(kFull, pv, base + 1);
  const int64_t e0 = __shfl_sync(kFull, pv, base + 2), e1 = __shfl_sync(kFull, pv, base + 3);
  const int64_t l0 = __shfl_sync(kFull, pv, base + 4), l1 = __shfl_sync(kFull, pv, base + 5);
  const bool fits = active && n1 >= n0 && e1 >= e0 && (n1 - n0) <= kMaxN && (e1 - e0) <= 32 * kEC &&
                    n0 >= nb0 && (n1 - nb0) <= kSlabNodes;
  const int n = fits ? static_cast<int>(n1 - n0) : 0;
  const int ne = fits ? static_cast<int>(e1 - e0) : 0;
  if (g == B - 1 && sl == 0) n_lut[0] = static_cast<int32_t>(l1);

  // ---- node slab -> shared memory (4-byte cp.async: slabs are only 4-byte aligned)
  float* sx = sm.x + (fits ? static_cast<int>(n0 - nb0) : 0) * kF;
  {
    const float* __restrict__ xg = x + n0 * kF;
    const int nf = n * kF;
    for (int i = sl; i < nf; i += 8) cp_async4(sx + i, xg + i);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- destination row: lane sl holds edges sl + 8j, j = 0..31, one graph-local byte each.  Only the
  // low words are fetched (ids are < N < 2^31 for every batch a collate or qot_edge_ptr vouches for;
  // the value is only ever compared, never used as an address), 16 loads in flight per lane.
  unsigned pk[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) pk[q] = 0xffffffffu;
  {
    const unsigned* __restrict__ dg = reinterpret_cast<const unsigned*>(edst + e0);
    const unsigned n0u = static_cast<unsigned>(n0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (__any_sync(kFull, 128 * half < ne)) {
        unsigned d[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int e = sl + 8 * (16 * half + t);
          d[t] = e < ne ? __ldg(dg + 2 * e) : 0xffffffffu;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const unsigned loc = d[t] - n0u;
          const unsigned byte = (e0 >= 0 && loc < static_cast<unsigned>(n) && sl + 8 * (16 * half + t) < ne) ? loc : 0xffu;
          const int q = 4 * half + (t >> 2);
          pk[q] = (pk[q] & ~(0xffu << (8 * (t & 3)))) | (byte << (8 * (t & 3)));
        }
      }
    }
  }
  cp_async_wait_all();
  __syncwarp();                                      // this sub-group's slab is in place

  // ---- LUT node(s) of the graph
  int cnt = 0, il = -1;
#pragma unroll
  for (int r = 0; r < kMaxN / 8; ++r) {
    const int node = sl + 8 * r;
    const unsigned bal = __ballot_sync(kFull, node < n && sx[node * kF + lut_col] == 1.0f);
    const unsigned sub = (bal >> (8 * sg)) & 0xffu;
    if (sub) {
      if (il < 0) il = 8 * r + __ffs(sub) - 1;
      cnt += __popc(sub);
    }
  }
  bool ok = fits && cnt == 1 && (l1 - l0) == 1;      // fast row: exactly one LUT node, as lut_ptr says
  if (fits && sl == 0 && cnt != l1 - l0) atomicOr(status, 1);   // lut_ptr does not describe this x

  // ---- in-edges of the LUT node: SIMD byte compare, then compaction in edge order
  int* msg = sm.msg[gl];
  int mc = 0;
  {
    const unsigned pat = static_cast<unsigned>(il & 0xff) * 0x01010101u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const unsigned m4 = ok ? __vcmpeq4(pk[q], pat) : 0u;
      if (__any_sync(kFull, m4 != 0u)) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool hit = (m4 >> (8 * t)) & 1u;
          const unsigned sub = (__ballot_sync(kFull, hit) >> (8 * sg)) & 0xffu;
          const int pos = mc + __popc(sub & ((1u << sl) - 1u));
          if (hit && pos < kSubMsg - 1) msg[pos] = sl + 8 * (4 * q + t);
          mc += __popc(sub);
        }
      }
    }
  }
  if (mc > kSubMsg - 1) ok = false;                  // hub row: generic path
  __syncwarp();
  // sources of those edges (one gather per 8); self loops / out-of-range ids dropped, order kept
  {
    int kept = 0;
    bool outside = false;
#pragma unroll
    for (int t0 = 0; t0 < kSubMsg; t0 += 8) {
      if (__any_sync(kFull, ok && t0 < mc)) {
        const int t = t0 + sl;
        long long sj = -1;
        if (ok && t < mc) sj = esrc[e0 + msg[t]];
        const bool inN = static_cast<uint64_t>(sj) < static_cast<uint64_t>(N);
        const long long sloc = sj - n0;
        const bool inslab = sloc >= 0 && sloc < n;
        const bool keep = inslab && sloc != il;
        outside |= ((__ballot_sync(kFull, inN && !inslab) >> (8 * sg)) & 0xffu) != 0u;
        const unsigned sub = (__ballot_sync(kFull, keep) >> (8 * sg)) & 0xffu;
        __syncwarp();
        if (keep) msg[kept + __popc(sub & ((1u << sl) - 1u))] = static_cast<int>(sloc);
        kept += __popc(sub);
        __syncwarp();
      }
    }
    if (outside) ok = false;                          // a source outside the slab: generic path
    mc = kept;
  }
  if (ok && sl == 0) msg[mc] = il;                    // the appended self loop comes last
  ++mc;
  __syncwarp();

  // ---- attention: the 8 lanes are 8 message slots, each lane carries all 4 heads
  {
    float As[kF][kHeads], d[kHeads];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) d[h] = 0.f;
    const int ils = ok ? il : 0;
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      const float xi = sx[ils * kF + k];
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        As[k][h] = __ldg(prep + kOffAsrc + k * kHeads + h);
        d[h] = fmaf(xi, __ldg(prep + kOffAdst + k * kHeads + h), d[h]);
      }
    }
    float mx[kHeads], ssum[kHeads], acc[kHeads][kF];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
      mx[h] = -INFINITY;
      ssum[h] = 0.f;
#pragma unroll
      for (int k = 0; k < kF; ++k) acc[h][k] = 0.f;
    }
#pragma unroll
    for (int r0 = 0; r0 < kSubMsg; r0 += 8) {
      if (__any_sync(kFull, ok && r0 < mc)) {
        const bool valid = ok && r0 + sl < mc;
        const int j = valid ? msg[r0 + sl] : 0;
        float xj[kF];
#pragma unroll
        for (int k = 0; k < kF; ++k) xj[k] = sx[j * kF + k];
#pragma unroll
        for (int h = 0; h < kHeads; ++h) {
          float a = d[h];
#pragma unroll
          for (int k = 0; k < kF; ++k) a = fmaf(xj[k], As[k][h], a);
          a = a > 0.f ? a : 0.2f * a;
          if (!valid) a = -INFINITY;
          float mr = a;
#pragma unroll
          for (int o = 1; o <= 4; o <<= 1) mr = fmaxf(mr, __shfl_xor_sync(kFull, mr, o));
          const float mn = fmaxf(fmaxf(mx[h], mr), -1e30f);   // idle sub-groups stay finite
          const float sc = expf(mx[h] - mn);
          const float p = valid ? expf(a - mn) : 0.f;
          float v[kF + 1];
          v[kF] = p;
#pragma unroll
          for (int k = 0; k < kF; ++k) v[k] = p * xj[k];
#pragma unroll
          for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
            for (int k = 0; k <= kF; ++k) v[k] += __shfl_xor_sync(kFull, v[k], o);
          }
          ssum[h] = fmaf(ssum[h], sc, v[kF]);
#pragma unroll
          for (int k = 0; k < kF; ++k) acc[h][k] = fmaf(acc[h][k], sc, v[k]);
          mx[h] = mn;
        }
      }
    }
    if (ok && sl < kF) {
#pragma unroll
      for (int h = 0; h < kHeads; ++h) sm.z[gl][h * kF + sl] = pick5(acc[h], sl) / (ssum[h] + 1e-16f);
    }
  }
  if (sl == 0) {
    SubMeta& mt = sm.meta[gl];
    mt.n0 = n0; mt.n1 = n1; mt.e0 = e0; mt.e1 = e1; mt.l0 = l0; mt.l1 = l1;
    mt.il = il;
    // generic path: anything active that is not a finished fast row and may own LUT rows
    mt.state = ok ? 1 : (active && (!fits || (cnt == l1 - l0 && cnt > 0)) ? 2 : 0);
  }
  __syncthreads();

  // ---- readout heads, two rows per warp at a time; the 19.5 KB of weights are read through L1
  const float* __restrict__ wh = prep + kOffWf;
#pragma unroll 1
  for (int p = 0; p < 2; ++p) {
    const int ga = warp + 16 * p, gb = ga + 8;
    const bool sa = sm.meta[ga].state == 1, sb = sm.meta[gb].state == 1;
    if (!sa && !sb) continue;
    float oa, ob;
    lut_head2(wh, sm.z[sa ? ga : gb], sm.z[sb ? gb : ga], sm.y[warp][0], sm.y[warp][1], lane, oa, ob);
    __syncwarp();
    if (sa) {
      const SubMeta& mt = sm.meta[ga];
      if (lane < QOT_OUT) out[mt.l0 * QOT_OUT + lane] = oa;
      if (lane == 0) {
        lut_batch[mt.l0] = static_cast<int64_t>(blockIdx.x) * kGPB + ga;
        lut_node[mt.l0] = static_cast<int32_t>(mt.n0 + mt.il);
      }
    }
    if (sb) {
      const SubMeta& mt = sm.meta[gb];
      if (lane < QOT_OUT) out[mt.l0 * QOT_OUT + lane] = ob;
      if (lane == 0) {
        lut_batch[mt.l0] = static_cast<int64_t>(blockIdx.x) * kGPB + gb;
        lut_node[mt.l0] = static_cast<int32_t>(mt.n0 + mt.il);
      }
    }
  }

  // ---- generic path for the graphs the fast path declined (one warp per graph)
  for (int gs = warp; gs < kGPB; gs += 8) {
    if (sm.meta[gs].state != 2) continue;
    const SubMeta mt = sm.meta[gs];
    const int64_t gg = static_cast<int64_t>(blockIdx.x) * kGPB + gs;
    float* s_y = sm.y[warp][0];
    float* s_z = sm.y[warp][1];
    int* s_m = reinterpret_cast<int*>(sm.y[warp][1] + 32);
    int64_t orow = mt.l0;
    int found = 0;
    for (int64_t nb = mt.n0; nb < mt.n1; nb += 32) {
      const int64_t node = nb + lane;
      unsigned mask = __ballot_sync(kFull, node < mt.n1 && x[node * kF + lut_col] == 1.0f);
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        ++found;
        if (orow < mt.l1) {                                      // never write past this graph's rows
          const int64_t i = nb + bit;
          const float ov = lut_row_global(x, esrc, edst, mt.e0, mt.e1, N, i, prep, wh, s_m, s_z, s_y, lane);
          if (lane < QOT_OUT) out[orow * QOT_OUT + lane] = ov;
          if (lane == 0) {
            lut_batch[orow] = gg;
            lut_node[orow] = static_cast<int32_t>(i);
          }
          ++orow;
        }
      }
    }
    if (lane == 0 && found != mt.l1 - mt.l0) atomicOr(status, 1);
  }
}

int lp_infer_launch_legacy(int variant, const float* x, const int64_t* esrc, const int64_t* edst, const int64_t* gptr,
                           const int64_t* eptr, const int64_t* lut_ptr, int64_t N, int64_t B, const float* prepared,
                           int32_t is_lut_index, float* out, int64_t* lut_batch, int32_t* lut_node, int32_t* n_lut,
                           int32_t* status, cudaStream_t stream) {
  if (variant == 1) {
    static bool attr_set = false;
    if (!attr_set) {
      QOT_CUDA(cudaFuncSetAttribute(lp_infer_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(SubSmem))));
      attr_set = true;
    }
    const int64_t blocks = cdiv(B, kGPB);
    lp_infer_sub_kernel<<<static_cast<unsigned>(blocks), 256, sizeof(SubSmem), stream>>>(
        x, esrc, edst, gptr, eptr, lut_ptr, N, B, prepared, is_lut_index, out, lut_batch, lut_node, n_lut, status);
    QOT_LAUNCH_CHECK();
    return QOT_OK;
  }
  const int64_t blocks = cdiv(B, kIW);
  QOT_REQUIRE(blocks < (1ll << 31) - 1, "qot_lightpath_infer: too many graphs for one launch");
  lp_infer_kernel<<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(
      x, esrc, edst, gptr, eptr, lut_ptr, N, B, prepared, is_lut_index, out, lut_batch, lut_node, n_lut, status);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

}  // namespace qot
