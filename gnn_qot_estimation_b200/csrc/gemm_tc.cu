// Node-wise dense projection on the 5th-generation tensor cores with fp32-level accuracy:
//   C[M,Nc] = A[rows(M),K] * W[Nc,K]^T (+ bias),   fp32 in, fp32 out,
// computed as three TF32 tcgen05.mma products per k-step on error-compensated operands
//   a = a_hi + a_lo,  w = w_hi + w_lo   (hi = round-to-tf32, lo = tf32(a - hi)),
//   a*w ~= a_hi*w_lo + a_lo*w_hi + a_hi*w_hi      (the dropped a_lo*w_lo term is ~2^-22 relative)
// accumulated in fp32 -- short chains in TMEM, the chains summed round-to-nearest in registers.  This is the one
// place on the path where the hidden width makes the projection a real contraction (BASELINE cfg 5:
// [10000,256] x [256,1024] and x [256,2560], 18 GFLOP); at the reference widths (H = 16 / 32) the FFMA kernel in
// dense.cu stays in charge.
//
// Two steps.  (1) split_tf32_kernel (and its transposing variant for weight gradients) writes the hi / lo halves of
// an operand once -- performing the embedding-row gather of topological_training/models.py:52 on the way -- as a
// sequence of 16 KB TILE IMAGES: 128 rows x 32 columns in exactly the 128-byte-swizzled K-major shared-memory layout
// the UMMA descriptors name.  (2) gemm_tf32x3_kernel: persistent, clusters of two CTAs, 128 x 128 output tiles.
// Per CTA one thread fetches tile images with bulk copies (cp.async.bulk: three per k-block, one of them multicast
// to both CTAs of the cluster) through three 64 KB stages on full / empty mbarriers, one thread issues the 12 MMAs of
// a k-block, and four warps drain: a TMEM accumulator (one of four, 128 lanes x 128 columns) only ever holds a chain
// of TC_CHAIN k-blocks, which the drain warps read with tcgen05.ld and add into registers; epilogue: registers + bias
// -> global while the next tile is already in the tensor pipe.  profiles/r2_cfg5_forward.md has the measurements
// behind each of these choices.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;          // 32 fp32 = one 128-byte swizzle row
#ifndef QOT_TC_STAGES
#define QOT_TC_STAGES 3
#endif
constexpr int TC_STAGES = QOT_TC_STAGES;
constexpr int TC_THREADS = 192;                               // warp 0: bulk-copy producer, 1: MMA issue, 2-5: drain + epilogue
#ifndef QOT_TC_CHAIN
#define QOT_TC_CHAIN 2
#endif
constexpr int TC_ACCS = 4;                                    // TMEM accumulators (4 x 128 columns = all of TMEM), used round-robin by the chains
constexpr int64_t TC_CHAIN = QOT_TC_CHAIN;                     // k-blocks per TMEM accumulation chain (see the main loop)
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;            // 16 KB per (operand, hi|lo) tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;           // A_hi, A_lo, W_hi, W_lo
constexpr int TC_EPI_PITCH = 36;                              // floats per row of an epilogue patch
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers, TMEM slot*/ +
                              4 * 32 * TC_EPI_PITCH * 4 /*epilogue patches*/;
constexpr unsigned TC_SPIN_LIMIT = 1u << 26;                // a wedged barrier ends the kernel, never hangs it

// -DQOT_TC_TRACE: per-role cycle accounting (scripts/trace_gemm_tc.py): 8 counters per CTA
#ifdef QOT_TC_TRACE
__device__ unsigned long long* g_tc_trace = nullptr;
#define TC_T() const long long tc_t_ = clock64()
#define TC_ACC(slot) do { tc_acc_[slot] += clock64() - tc_t_; } while (0)
#define TC_DECL() long long tc_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TC_FLUSH() do { if (g_tc_trace) for (int i_ = 0; i_ < 8; ++i_) if (tc_acc_[i_]) atomicAdd(g_tc_trace + blockIdx.x * 8 + i_, static_cast<unsigned long long>(tc_acc_[i_])); } while (0)
#else
#define TC_T() do {} while (0)
#define TC_ACC(slot) do {} while (0)
#define TC_DECL() do {} while (0)
#define TC_FLUSH() do {} while (0)
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
// explicit shared-space accesses: the patch pointer is derived from an aligned-up generic address, so plain
// dereferences compile to generic LD / ST (long-scoreboard latency) instead of LDS / STS
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float to_tf32(float v) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address,
// LBO (unused for a 128-byte-wide K extent) = 1, SBO = 1024 B (8 rows x 128 B), version 1, layout 2.
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= static_cast<unsigned long long>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<unsigned long long>(1u) << 16;
  d |= static_cast<unsigned long long>(1024u >> 4) << 32;
  d |= static_cast<unsigned long long>(1u) << 46;
  d |= static_cast<unsigned long long>(2u) << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N, M
constexpr unsigned tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<unsigned>(N >> 3) << 17) |
         (static_cast<unsigned>(M >> 4) << 24);
}
// The three products of one k-step (UMMA_K = 8 columns = 32 bytes of the swizzled rows) in one block:
// D (+)= A_hi W_lo^T; D += A_lo W_hi^T; D += A_hi W_hi^T -- small terms first.  The descriptors differ from a per-stage
// base only in the start-address field ((addr >> 4) & 0x3fff, the low bits), so the issuing thread forms them by
// ADDING to the base (the one-thread issue path, not the tensor pipe, was the kernel's cycle time when every
// descriptor was rebuilt from its address: 75 cycles per MMA, profiles/r2_cfg5_forward.md).
__device__ __forceinline__ void umma_tf32_step(unsigned tmem_d, unsigned long long a_hi, unsigned long long a_lo,
                                               unsigned long long w_hi, unsigned long long w_lo, unsigned idesc,
                                               unsigned accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %5, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %5, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %5, t;\n\t}\n" ::"r"(tmem_d),
      "l"(a_hi), "l"(a_lo), "l"(w_hi), "l"(w_lo), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spin = 0; spin < TC_SPIN_LIMIT; ++spin) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
// ---- operand images.  The pre-pass writes the hi / lo halves of an operand as a sequence of TILES: tile (rb, kb)
// holds rows [128 rb, 128 rb + 128) x columns [32 kb, 32 kb + 32) as 16 KB in exactly the shared-memory image the
// UMMA descriptor names (row r at r * 128 B, its 16-byte chunk j stored at chunk j ^ (r & 7): SWIZZLE_128B), tiles
// ordered kb fastest.  The GEMM then fetches a tile with ONE bulk copy (cp.async.bulk, UBLKCP) -- no per-thread
// address arithmetic, no tensor map.  Rows past the matrix edge are written as zeros (the image is padded to 128 rows).
__device__ __forceinline__ int64_t tc_img_offset(int64_t r, int64_t chunk /* 16-byte chunk index along K */, int64_t KB) {
  const int64_t rb = r >> 7, rr = r & 127, kb = chunk >> 3, j = chunk & 7;
  return ((rb * KB + kb) * TC_BM + rr) * TC_BK + ((j ^ (rr & 7)) << 2);
}

// hi = tf32(v), lo = tf32(v - hi) of rows [0, rows) of `src` (optional gather), K columns; rows [rows, rows_pad) zero.
// One warp per row per pass, four rows in flight per warp (independent 128-bit loads).
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ gather,
                  int64_t rows, int64_t rows_pad, int64_t K, float* __restrict__ hi, float* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const int kv = static_cast<int>(K / 4);
  const int64_t KB = K / TC_BK;
  const int64_t r0 = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * 4;
  for (int c = lane; c < kv; c += 32) {
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v[i] = *reinterpret_cast<const float4*>(src + (gather ? gather[r] : r) * ld + c * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + i;
      if (r < rows_pad) {
        float4 h, l;
        h.x = to_tf32(v[i].x); h.y = to_tf32(v[i].y); h.z = to_tf32(v[i].z); h.w = to_tf32(v[i].w);
        l.x = to_tf32(v[i].x - h.x); l.y = to_tf32(v[i].y - h.y); l.z = to_tf32(v[i].z - h.z); l.w = to_tf32(v[i].w - h.w);
        const int64_t o = tc_img_offset(r, c, KB);
        *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = l;
      }
    }
  }
}

// Transposing variant for weight gradients: src [R, Cc] (optional row gather) -> images of the [Cc, Rpad] matrix,
// i.e. the reduction dimension R becomes the contiguous (K-major) one; entries with r >= R or c >= Cc are zero
// (Cpad = Cc rounded up to 128 rows, Rpad to 32 columns).
__global__ void __launch_bounds__(256)
split_tf32_transpose_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ gather,
                            int64_t R, int64_t Cc, int64_t Rpad, int64_t Cpad, float* __restrict__ hi,
                            float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32, c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int64_t KB = Rpad / TC_BK;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) v = src[(gather ? gather[r] : r) * ld + c];
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < Cpad && r < Rpad) {
      const float v = tile[tx][ty + 8 * i];
      const float h = to_tf32(v);
      const int64_t o = tc_img_offset(c, r >> 2, KB) + (r & 3);
      hi[o] = h;
      lo[o] = to_tf32(v - h);
    }
  }
}
// out[i] = sum_z part[z*n + i], z ascending (fixed order)
__global__ void tc_reduce_splits_kernel(const float* __restrict__ part, int64_t n, int splits,
                                        float* __restrict__ out, int64_t ncols, int64_t ld_out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int z = 0; z < splits; ++z) a += part[static_cast<int64_t>(z) * n + i];
  out[(i / ncols) * ld_out + (i % ncols)] = a;
}

// One tile of the launch: output block (m0, n0) and, for split-K, the k-block slice [kb0, kb0 + nkb) and the
// partial output it is written to.  Tiles are numbered n-block fastest, so the CTAs working side by side share a few
// row blocks of A and all of W in L2.
// Work is handed out in PAIRS of tiles to clusters of two CTAs.  The two tiles of a pair share one operand block --
// the same rows of A (pair along n, `pair_n`) or the same rows of W (pair along m) -- and each CTA fetches only HALF
// of the shared operand (rank 0 the hi image, rank 1 the lo image) and multicasts it into both CTAs' shared memory:
// 48 KB instead of 64 KB through L2 per CTA and k-block.  A pair whose second tile falls off the matrix edge still
// runs it (clamped onto the first, not stored): the partner needs its half of the shared operand.
struct TcTile {
  int64_t m0, n0, kb0, nkb;
  float* C;
  bool store;
};
struct TcGrid {
  int64_t tiles_m, tiles_n, kb_total, kb_per_split, splits, npairs;
  bool pair_n;
};
__host__ __device__ inline TcGrid tc_grid(int64_t M, int64_t Nc, int64_t K, int64_t kb_per_split, int64_t splits) {
  TcGrid g;
  g.tiles_m = (M + TC_BM - 1) / TC_BM;
  g.tiles_n = (Nc + TC_BN - 1) / TC_BN;
  g.kb_total = K / TC_BK;
  g.kb_per_split = kb_per_split;
  g.splits = splits;
  g.pair_n = g.tiles_n % 2 == 0 || g.tiles_m % 2 != 0;              // prefer the dimension with an even tile count
  g.npairs = splits * (g.pair_n ? g.tiles_m * ((g.tiles_n + 1) / 2) : ((g.tiles_m + 1) / 2) * g.tiles_n);
  return g;
}
__host__ __device__ inline TcTile tc_tile(const TcGrid& g, int64_t p, int rank, float* C, int64_t M, int64_t Nc) {
  TcTile ti;
  int64_t z, mb, nb;
  if (g.pair_n) {
    const int64_t pn = (g.tiles_n + 1) / 2, per = g.tiles_m * pn, r = p % per;
    z = p / per; mb = r / pn; nb = 2 * (r % pn) + rank;
    ti.store = nb < g.tiles_n;
    nb = nb < g.tiles_n ? nb : g.tiles_n - 1;
  } else {
    const int64_t pm = (g.tiles_m + 1) / 2, per = pm * g.tiles_n, r = p % per;
    z = p / per; mb = 2 * (r / g.tiles_n) + rank; nb = r % g.tiles_n;
    ti.store = mb < g.tiles_m;
    mb = mb < g.tiles_m ? mb : g.tiles_m - 1;
  }
  ti.m0 = mb * TC_BM;
  ti.n0 = nb * TC_BN;
  ti.kb0 = z * g.kb_per_split;
  const int64_t left = g.kb_total - ti.kb0;
  ti.nkb = left < 0 ? 0 : (left < g.kb_per_split ? left : g.kb_per_split);
  ti.C = g.splits > 1 ? C + z * M * Nc : C;
  return ti;
}

// Persistent, clusters of two CTAs: cluster c works on tile pairs c, c + #clusters, ...  Warp roles in each CTA:
//   warp 0 (one lane)  producer: per k-block three bulk copies of 16 KB tile images into one of three 64 KB stages --
//                      the two images of its own operand block and ONE image of the shared block, multicast to both
//                      CTAs -- completing on the stage's `full` mbarrier (which also receives the partner's
//                      multicast); it runs straight across tile boundaries, so the next tile's first blocks are in
//                      flight while this tile's last MMAs run;
//   warp 1 (one lane)  MMA issue: waits `full`, issues the 12 tcgen05.mma of the block, commits them to the stage's
//                      `empty` barrier of BOTH CTAs (the partner writes into this stage too) and, at the end of a
//                      chain, to the chain barrier;
//   warps 2-5          drain + epilogue (TMEM lane quarter = warp % 4): add each finished TMEM chain into registers,
//                      write the tile when its last chain is in, while the next tile is already being multiplied.
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32x3_kernel(const float* __restrict__ Ahi, const float* __restrict__ Alo,
                   const float* __restrict__ Whi, const float* __restrict__ Wlo,
                   const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int64_t M,
                   int64_t Nc, int64_t K, int64_t kb_per_split, int64_t splits, int32_t* __restrict__ status) {
  extern __shared__ char tc_smem_raw[];
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = static_cast<int>(rank_u);
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + TC_STAGES * TC_STAGE_BYTES);
  unsigned long long* empty = full + TC_STAGES;                   // [STAGES] MMAs that read the stage have finished
  unsigned long long* chain_full = empty + TC_STAGES;             // [ACCS] MMA thread -> drain warps: chain finished
  unsigned long long* chain_free = chain_full + TC_ACCS;          // [ACCS] drain warps -> MMA thread: accumulator read out
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(chain_free + TC_ACCS);
  float* epi = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES + 128);   // 4 warps x 32 x 36 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcGrid tg = tc_grid(M, Nc, K, kb_per_split, splits);
  const int64_t npairs = tg.npairs, cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  auto tile = [&](int64_t pair) { return tc_tile(tg, pair, rank, C, M, Nc); };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<unsigned>(TC_ACCS * TC_BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(empty + s)) : "memory");   // both CTAs' MMAs
    }
    for (int s = 0; s < TC_ACCS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(chain_full + s)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(chain_free + s)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // both CTAs' barriers exist before either multicasts into / arrives on the other's
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_d = *tmem_slot;
  constexpr unsigned idesc = tc_idesc(TC_BM, TC_BN);
  bool ok = true;

  // Accumulation happens in two places.  The tensor core adds into TMEM with truncation, so the error of a chain of
  // MMAs on one accumulator grows LINEARLY with its length (profiles/r2_tc_chain_accuracy.md: max error / max|C| of
  // [10000,256]x[256,1024] is 3.1e-7, 5.5e-7, 1.1e-6, 2.5e-6 for chains of 1, 2, 4, 8 k-blocks; a weight gradient
  // reduced over 80 000 rows in one chain per split-K slice sat at 9e-6, torch's fp32 GEMM at 8e-7).  So a TMEM
  // accumulator only ever holds a CHAIN of TC_CHAIN k-blocks (2: 24 MMAs).  Four accumulators take turns (chains are
  // numbered through the whole tile sequence of the CTA; the MMA thread can run three chains ahead of the drain); the drain warpgroup (TMEM lane quarter = warp % 4) reads
  // each finished chain with tcgen05.ld and adds it, round-to-nearest, into fp32 registers (128 per thread: its row of
  // the tile), off the load -> MMA -> refill critical path.
  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      const int64_t KB = tg.kb_total;
      TC_DECL();
      int64_t g = 0;                                                   // blocks fetched so far (stage = g % STAGES)
      for (int64_t t = cluster; t < npairs; t += nclusters) {
        const TcTile ti = tile(t);
        const int64_t mb = ti.m0 / TC_BM, nb = ti.n0 / TC_BN;
        for (int64_t kb = 0; kb < ti.nkb; ++kb, ++g) {
          const int st = static_cast<int>(g % TC_STAGES);
          {
            TC_T();
            if (g >= TC_STAGES)                                        // both CTAs are done reading the stage
              ok &= mbar_wait(smem_u32(empty + st), static_cast<unsigned>(((g / TC_STAGES) - 1) & 1));
            TC_ACC(0);
          }
          TC_T();
          const unsigned fb = smem_u32(full + st);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(TC_STAGE_BYTES) : "memory");
          const unsigned dst = smem_u32(smem + st * TC_STAGE_BYTES);   // [A_hi | A_lo | W_hi | W_lo]
          const int64_t ao = (mb * KB + ti.kb0 + kb) * (TC_BM * TC_BK), wo = (nb * KB + ti.kb0 + kb) * (TC_BN * TC_BK);
          // own operand block: both images, this CTA only
          const float* own_hi = tg.pair_n ? Whi + wo : Ahi + ao;
          const float* own_lo = tg.pair_n ? Wlo + wo : Alo + ao;
          const unsigned own_dst = dst + (tg.pair_n ? 2 * TC_TILE_BYTES : 0);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(own_dst),
                       "l"(own_hi), "r"(TC_TILE_BYTES), "r"(fb) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           own_dst + TC_TILE_BYTES), "l"(own_lo), "r"(TC_TILE_BYTES), "r"(fb) : "memory");
          // shared operand block: rank 0 fetches the hi image, rank 1 the lo image, each for both CTAs
          const float* sh = tg.pair_n ? (rank == 0 ? Ahi + ao : Alo + ao) : (rank == 0 ? Whi + wo : Wlo + wo);
          const unsigned sh_dst = dst + (tg.pair_n ? 0 : 2 * TC_TILE_BYTES) + rank * TC_TILE_BYTES;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                  sh_dst), "l"(sh), "r"(TC_TILE_BYTES), "r"(fb), "h"(static_cast<unsigned short>(3)) : "memory");
          TC_ACC(1);
        }
      }
      TC_FLUSH();
    }
  } else if (warp == 1) {
    // ================= MMA issue =================
    if (lane == 0) {
      int64_t g = 0, gc = 0;                                           // blocks multiplied / chains started so far
      const unsigned long long desc0 = umma_desc_sw128(smem_u32(smem));   // stage 0, A_hi, k-step 0
      TC_DECL();
      for (int64_t t = cluster; t < npairs; t += nclusters) {
        const TcTile ti = tile(t);
        for (int64_t kb = 0; kb < ti.nkb; ++kb, ++g) {
          const int st = static_cast<int>(g % TC_STAGES);
          const int64_t c = gc + kb / TC_CHAIN;                        // chain of this block, accumulator c % ACCS
          const int ac = static_cast<int>(c % TC_ACCS);
          {
            TC_T();
            if (kb % TC_CHAIN == 0 && c >= TC_ACCS)                    // chain c-ACCS must have been read out
              ok &= mbar_wait(smem_u32(chain_free + ac), static_cast<unsigned>(((c / TC_ACCS) - 1) & 1));
            TC_ACC(2);
          }
          {
            TC_T();
            ok &= mbar_wait(smem_u32(full + st), static_cast<unsigned>((g / TC_STAGES) & 1));
            TC_ACC(3);
          }
          TC_T();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // stage image [A_hi | A_lo | W_hi | W_lo]: descriptor = stage-0 base + (byte offset >> 4)
          const unsigned long long a_hi = desc0 + static_cast<unsigned>(st * (TC_STAGE_BYTES >> 4));
          const unsigned long long a_lo = a_hi + (TC_TILE_BYTES >> 4), w_hi = a_hi + 2 * (TC_TILE_BYTES >> 4),
                                   w_lo = a_hi + 3 * (TC_TILE_BYTES >> 4);
          const unsigned d = tmem_d + static_cast<unsigned>(ac * TC_BN);
          const bool fresh = kb % TC_CHAIN == 0;                       // first block of a chain overwrites
#pragma unroll
          for (int s = 0; s < TC_BK / 8; ++s)                          // UMMA_K = 8 for tf32: 32 bytes (2 x 16) per step
            umma_tf32_step(d, a_hi + 2 * s, a_lo + 2 * s, w_hi + 2 * s, w_lo + 2 * s, idesc, !(fresh && s == 0));
          // arrives on the stage's `empty` barrier of both CTAs when every MMA issued so far has finished reading
          // shared memory
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           smem_u32(empty + st)), "h"(static_cast<unsigned short>(3))
                       : "memory");
          if (kb % TC_CHAIN == TC_CHAIN - 1 || kb == ti.nkb - 1)      // ... and on the chain barrier: accumulator complete
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(chain_full + ac))
                         : "memory");
          TC_ACC(4);
        }
        gc += (ti.nkb + TC_CHAIN - 1) / TC_CHAIN;
      }
      TC_FLUSH();
    }
  } else {
    // ================= drain + epilogue warpgroup =================
    const int q = warp & 3;                                            // TMEM lanes [32q, 32q+32) = tile rows
    int64_t gc = 0;
    TC_DECL();
    for (int64_t mt = cluster; mt < npairs; mt += nclusters) {
      const TcTile ti = tile(mt);
      float acc[TC_BN];
#pragma unroll
      for (int j = 0; j < TC_BN; ++j) acc[j] = 0.f;
      const int64_t nchains = (ti.nkb + TC_CHAIN - 1) / TC_CHAIN;
      for (int64_t cc = 0; cc < nchains; ++cc, ++gc) {
        const int ac = static_cast<int>(gc % TC_ACCS);
        {
          TC_T();
          ok &= mbar_wait(smem_u32(chain_full + ac), static_cast<unsigned>((gc / TC_ACCS) & 1));
          TC_ACC(5);
        }
        TC_T();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned base = tmem_d + (static_cast<unsigned>(q * 32) << 16) + static_cast<unsigned>(ac * TC_BN);
#pragma unroll
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
          unsigned r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(base + static_cast<unsigned>(c0))
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(chain_free + ac)) : "memory");
        TC_ACC(6);
      }
      TC_T();
      // ---- epilogue: the thread holds ROW q*32+lane of the tile (an empty k-slice: zeros).  Written as it stands,
      // one store instruction would touch 32 rows x 16 bytes (7 300 cycles per tile, as long as the whole mainloop);
      // so each 32 x 32 chunk is turned through a per-warp shared-memory patch (pitch 36 floats: conflict-free both
      // ways) and leaves as 8 instructions of 4 rows x 128 contiguous bytes.
      if (ti.store) {
        const unsigned patch = smem_u32(epi) + static_cast<unsigned>(q * (32 * TC_EPI_PITCH * 4));   // byte address
        const int pr = lane >> 3, pc = (lane & 7) * 4;                 // read side: row within a group of 4, column
#pragma unroll
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            sts128(patch + static_cast<unsigned>((lane * TC_EPI_PITCH + j) * 4),
                   make_float4(acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]));
          __syncwarp();
          const int64_t col = ti.n0 + c0 + pc;
          const bool vec = (ldc & 3) == 0 && col + 4 <= Nc;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) {
            if (vec) bv = *reinterpret_cast<const float4*>(bias + col);
            else {
              bv.x = col < Nc ? bias[col] : 0.f; bv.y = col + 1 < Nc ? bias[col + 1] : 0.f;
              bv.z = col + 2 < Nc ? bias[col + 2] : 0.f; bv.w = col + 3 < Nc ? bias[col + 3] : 0.f;
            }
          }
#pragma unroll 2
          for (int k = 0; k < 8; ++k) {
            const int r = 4 * k + pr;
            const int64_t row = ti.m0 + q * 32 + r;
            float4 o = lds128(patch + static_cast<unsigned>((r * TC_EPI_PITCH + pc) * 4));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            if (row < M) {
              float* cp = ti.C + row * ldc + col;
              if (vec) *reinterpret_cast<float4*>(cp) = o;
              else {
                if (col < Nc) cp[0] = o.x;
                if (col + 1 < Nc) cp[1] = o.y;
                if (col + 2 < Nc) cp[2] = o.z;
                if (col + 3 < Nc) cp[3] = o.w;
              }
            }
          }
          __syncwarp();
        }
      }
      TC_ACC(7);
    }
    if (warp == 2 && lane == 0) TC_FLUSH();
  }
  if (!ok && status) atomicOr(status, 2);
  // the drain warpgroup leaves its loop only after the last chain barrier, i.e. after every MMA has finished; the
  // cluster barrier keeps this CTA's shared memory alive until the partner has stopped multicasting into it /
  // arriving on its barriers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(static_cast<unsigned>(TC_ACCS * TC_BN))
                 : "memory");
}

}  // namespace qot

using namespace qot;

// C[M,Nc] (ldc) = A[gather? gather[m] : m, :K] (lda) * W[Nc,:K]^T (ldw) (+ bias[Nc]); K % 32 == 0,
// lda/ldw % 4 == 0 and 16-byte aligned bases (128-bit accesses).  ws: qot_gemm_tf32x3_workspace_bytes(M,Nc,K)
// bytes for the hi / lo halves of both operands.  status (optional, device int32): bit 1 is set if a
// pipeline barrier timed out (never observed; the result is then undefined, the kernel still ends).
static int tc_attr() {
  static std::atomic<unsigned long long> done{0};
  return once_per_device(done, [] {
    QOT_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    return static_cast<int>(QOT_OK);
  });
}

// Launch: clusters of two CTAs, as many clusters as the device holds at once (asked from the occupancy API once per
// device: GPC boundaries decide how many SM pairs there are), each looping over the tile pairs.
static int tc_launch(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, const float* bias,
                     float* C, int64_t ldc, int64_t M, int64_t Nc, int64_t K, int64_t kb_per_split, int64_t splits,
                     int32_t* status, cudaStream_t stream) {
  static std::atomic<int> max_clusters[64];
  int dev = 0;
  QOT_CUDA(cudaGetDevice(&dev));
  QOT_REQUIRE(dev >= 0 && dev < 64, "qot_gemm_tf32x3: device ordinal %d out of range", dev);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int mc = max_clusters[dev].load(std::memory_order_acquire);
  if (mc == 0) {
    cfg.gridDim = dim3(2 * kNumSMs);
    QOT_CUDA(cudaOccupancyMaxActiveClusters(&mc, gemm_tf32x3_kernel, &cfg));
    QOT_REQUIRE(mc > 0, "qot_gemm_tf32x3: no cluster of two CTAs fits on device %d", dev);
    max_clusters[dev].store(mc, std::memory_order_release);
  }
  const TcGrid g = tc_grid(M, Nc, K, kb_per_split, splits);
  cfg.gridDim = dim3(static_cast<unsigned>(2 * std::min<int64_t>(g.npairs, mc)));
  QOT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel, a_hi, a_lo, w_hi, w_lo, bias, C, ldc, M, Nc, K, kb_per_split,
                              splits, status));
  return QOT_OK;
}

static int64_t tc_pad128(int64_t n) { return cdiv(n, static_cast<int64_t>(TC_BM)) * TC_BM; }   // operand images are whole 128-row tiles

#ifdef QOT_TC_TRACE
extern "C" int qot_debug_set_tc_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -3;
}
#endif

// Test hook (host only, no GPU): the tile decomposition of a launch -- for every (pair, cluster rank) the output
// block, its k-block slice and whether the CTA stores it -- so that a CPU test can check that every output tile of
// every split-K slice is written exactly once.  out: [2 * npairs][6] = m0, n0, kb0, nkb, store, slice.  Returns the
// number of rows (2 * npairs); fills at most `cap` of them.
extern "C" int64_t qot_debug_gemm_tiles(int64_t M, int64_t Nc, int64_t K, int64_t kb_per_split, int64_t splits,
                                        int64_t* out, int64_t cap) {
  if (M <= 0 || Nc <= 0 || K <= 0 || K % TC_BK || kb_per_split <= 0 || splits <= 0) return -1;
  const TcGrid g = tc_grid(M, Nc, K, kb_per_split, splits);
  int64_t n = 0;
  for (int64_t p = 0; p < g.npairs; ++p)
    for (int rank = 0; rank < 2; ++rank, ++n) {
      if (!out || n >= cap) continue;
      const TcTile t = tc_tile(g, p, rank, nullptr, M, Nc);
      int64_t* o = out + n * 6;
      o[0] = t.m0; o[1] = t.n0; o[2] = t.kb0; o[3] = t.nkb; o[4] = t.store ? 1 : 0; o[5] = t.kb0 / kb_per_split;
    }
  return n;
}

extern "C" size_t qot_gemm_tf32x3_workspace_bytes(int64_t M, int64_t Nc, int64_t K) {
  if (M < 0 || Nc < 0 || K < 0) return 0;
  return 2 * align_up(static_cast<size_t>(tc_pad128(M)) * K * 4) + 2 * align_up(static_cast<size_t>(tc_pad128(Nc)) * K * 4) + 256;
}

extern "C" int qot_gemm_tf32x3(const float* A, int64_t lda, const int64_t* gather, const float* W, int64_t ldw,
                               const float* bias, float* C, int64_t ldc, int64_t M, int64_t Nc, int64_t K,
                               int32_t* status, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(M >= 0 && Nc >= 0 && K > 0, "qot_gemm_tf32x3: bad size");
  if (M == 0 || Nc == 0) return QOT_OK;
  QOT_REQUIRE(A && W && C, "qot_gemm_tf32x3: null operand");
  QOT_REQUIRE(K % TC_BK == 0, "qot_gemm_tf32x3: K must be a multiple of 32 (got %lld)", (long long)K);
  QOT_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldc >= Nc, "qot_gemm_tf32x3: leading dimensions must be multiples of 4");
  QOT_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
              "qot_gemm_tf32x3: operands must be 16-byte aligned");
  QOT_REQUIRE(ws && ws_bytes >= qot_gemm_tf32x3_workspace_bytes(M, Nc, K), "qot_gemm_tf32x3: workspace too small");
  if (int rc = tc_attr()) return rc;
  Carver c(ws);
  const int64_t Mp = tc_pad128(M), Np = tc_pad128(Nc);
  float* a_hi = c.take<float>(Mp * K);
  float* a_lo = c.take<float>(Mp * K);
  float* w_hi = c.take<float>(Np * K);
  float* w_lo = c.take<float>(Np * K);
  split_tf32_kernel<<<static_cast<unsigned>(Mp / 32), 256, 0, stream>>>(A, lda, gather, M, Mp, K, a_hi, a_lo);
  QOT_LAUNCH_CHECK();
  split_tf32_kernel<<<static_cast<unsigned>(Np / 32), 256, 0, stream>>>(W, ldw, nullptr, Nc, Np, K, w_hi, w_lo);
  QOT_LAUNCH_CHECK();
  return tc_launch(a_hi, a_lo, w_hi, w_lo, bias, C, ldc, M, Nc, K, K / TC_BK, 1, status, stream);
}

// Weight gradient on the tensor cores: C[Mo,No] (ldc) = sum_r A[r,:Mo]^T B[r,:No] over R rows (row-major
// A [R,lda], B [R,ldb]; optional row gather on B: the embedding lookup of the forward).  Both operands are
// transposed + hi/lo-split into K-major form (reduction over rows contiguous, padded to 32), the reduction is
// cut into split-K slices so the few output tiles still fill the machine, and the slices are summed in a fixed
// order -- deterministic.
static int tc_wgrad_splits(int64_t tiles, int64_t nkb) {
  // persistent kernel, one CTA per SM: as many slices as give every SM one (tile, slice) -- one even wave
  const int64_t want = std::max<int64_t>(1, kNumSMs / std::max<int64_t>(tiles, 1));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, nkb / 4 + 1), 64)));
}
static int64_t tc_rpad(int64_t R) { return cdiv(R, TC_BK) * TC_BK; }

extern "C" size_t qot_wgrad_tf32x3_workspace_bytes(int64_t R, int64_t Mo, int64_t No) {
  if (R < 0 || Mo < 0 || No < 0) return 0;
  const int64_t rp = tc_rpad(R);
  const int splits = tc_wgrad_splits(cdiv(Mo, TC_BM) * cdiv(No, TC_BN), rp / TC_BK);
  return 2 * align_up(static_cast<size_t>(tc_pad128(Mo)) * rp * 4) + 2 * align_up(static_cast<size_t>(tc_pad128(No)) * rp * 4) +
         align_up(static_cast<size_t>(splits) * Mo * No * 4) + 256;
}

extern "C" int qot_wgrad_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb,
                                const int64_t* gather_b, int64_t R, int64_t Mo, int64_t No, float* C,
                                int64_t ldc, int32_t* status, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(R > 0 && Mo > 0 && No > 0, "qot_wgrad_tf32x3: bad size");
  QOT_REQUIRE(A && B && C && ldc >= No, "qot_wgrad_tf32x3: null operand");
  QOT_REQUIRE(ws && ws_bytes >= qot_wgrad_tf32x3_workspace_bytes(R, Mo, No), "qot_wgrad_tf32x3: workspace too small");
  if (int rc = tc_attr()) return rc;
  const int64_t rp = tc_rpad(R), nkb = rp / TC_BK;
  const int64_t tiles = cdiv(Mo, TC_BM) * cdiv(No, TC_BN);
  const int splits = tc_wgrad_splits(tiles, nkb);
  const int64_t kps = cdiv(nkb, splits);
  Carver c(ws);
  const int64_t Mp = tc_pad128(Mo), Np = tc_pad128(No);
  float* at_hi = c.take<float>(Mp * rp);
  float* at_lo = c.take<float>(Mp * rp);
  float* bt_hi = c.take<float>(Np * rp);
  float* bt_lo = c.take<float>(Np * rp);
  float* part = c.take<float>(static_cast<size_t>(splits) * Mo * No);
  dim3 ga(static_cast<unsigned>(rp / 32), static_cast<unsigned>(Mp / 32));
  dim3 gb(static_cast<unsigned>(rp / 32), static_cast<unsigned>(Np / 32));
  QOT_REQUIRE(ga.y <= 65535u && gb.y <= 65535u, "qot_wgrad_tf32x3: operand too wide for one launch");
  split_tf32_transpose_kernel<<<ga, 256, 0, stream>>>(A, lda, nullptr, R, Mo, rp, Mp, at_hi, at_lo);
  QOT_LAUNCH_CHECK();
  split_tf32_transpose_kernel<<<gb, 256, 0, stream>>>(B, ldb, gather_b, R, No, rp, Np, bt_hi, bt_lo);
  QOT_LAUNCH_CHECK();
  if (splits == 1) return tc_launch(at_hi, at_lo, bt_hi, bt_lo, nullptr, C, ldc, Mo, No, rp, kps, 1, status, stream);
  if (int rc = tc_launch(at_hi, at_lo, bt_hi, bt_lo, nullptr, part, No, Mo, No, rp, kps, splits, status, stream)) return rc;
  const int64_t n = Mo * No;
  tc_reduce_splits_kernel<<<static_cast<unsigned>(cdiv(n, 256)), 256, 0, stream>>>(part, n, splits, C, No, ldc);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}
