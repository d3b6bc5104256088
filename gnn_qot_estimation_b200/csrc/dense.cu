// Node-wise dense pieces in exact FP32 (FFMA tiles): the projections inside
// TransformerConv / NNConv (PyG's addmm calls), their
// transposes for the backward, deterministic weight-gradient and bias reductions.
//
// At the reference widths (H = 16, 32) these are skinny products (K = 16): not a
// tensor-core contraction, and the 1e-5 parity bar rules out plain TF32 anyway.
#include <algorithm>

#include "common.cuh"

namespace qot {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;   // 256 threads, 4x4 micro-tile

// C[M,Nc] = sum_k A(m,k) B(k,n)  over k in this block's K-slice (blockIdx.z).
// A(m,k) = A[row(m)*a_rs + k*a_cs], row(m) = gather ? gather[m] : m.
// When gridDim.z > 1 every slice writes its partial to C + z*M*Nc (ldc == Nc).
__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const int64_t* __restrict__ gather,
            const float* __restrict__ B, int64_t b_rs, int64_t b_cs, const float* __restrict__ bias,
            float* __restrict__ C, int64_t ldc, int64_t M, int64_t Nc, int64_t K, int64_t k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM, n0 = static_cast<int64_t>(blockIdx.y) * BN;
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  if (gridDim.z > 1) C += static_cast<int64_t>(blockIdx.z) * M * Nc;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader roles: 256 threads fetch a 64x16 A tile and a 16x64 B tile (4 elements each).
  // Thread->element maps are chosen so the fastest-varying thread index follows the
  // unit-stride dimension of the operand.
  const bool a_k_contig = (a_cs == 1);
  const bool b_n_contig = (b_cs == 1);

  // thread -> element maps of the two tile loaders (4 elements each), fixed over the k loop
  int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * 256;                   // 0..1023
    if (a_k_contig) { a_k[it] = idx % BK; a_m[it] = idx / BK; } else { a_m[it] = idx % BM; a_k[it] = idx / BM; }
    if (b_n_contig) { b_n[it] = idx % BN; b_k[it] = idx / BN; } else { b_k[it] = idx % BK; b_n[it] = idx / BK; }
  }
  int64_t a_row[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t gm = m0 + a_m[it];
    a_row[it] = gm < M ? (gather ? gather[gm] : gm) : -1;
  }
  auto fetch = [&](int64_t k0, float (&av)[4], float (&bv)[4]) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t gk = k0 + a_k[it];
      av[it] = (a_row[it] >= 0 && gk < kend) ? A[a_row[it] * a_rs + gk * a_cs] : 0.f;
      const int64_t gn = n0 + b_n[it], gk2 = k0 + b_k[it];
      bv[it] = (gn < Nc && gk2 < kend) ? B[gk2 * b_rs + gn * b_cs] : 0.f;
    }
  };
  float av[4], bv[4];
  if (kbeg < kend) fetch(kbeg, av, bv);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      As[a_k[it]][a_m[it]] = av[it];
      Bs[b_k[it]][b_n[it]] = bv[it];
    }
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK, av, bv);       // next tile in flight while this one is consumed
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      const float a[TM] = {a4.x, a4.y, a4.z, a4.w};
      const float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t gn = n0 + tx * TN + j;
      if (gn < Nc) C[gm * ldc + gn] = acc[i][j] + ((bias && blockIdx.z == 0) ? bias[gn] : 0.f);
    }
  }
}

// out[i] = sum_s part[s*n + i]: the fixed-order second stage.  One warp per output element: lanes
// stride over the slices, then a shuffle tree (the order depends on `splits` only).
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int64_t n, int splits,
                       float* __restrict__ out, int64_t ld_out, int64_t ncols) {
  const int lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (i >= n) return;
  float s = 0.f;
  for (int k = lane; k < splits; k += 32) s += part[static_cast<int64_t>(k) * n + i];
  s = warp_sum(s);
  if (lane == 0) out[(i / ncols) * ld_out + (i % ncols)] = s;
}

constexpr int kColsumRows = 256;   // rows per first-stage block

__global__ void __launch_bounds__(256)
colsum_stage1_kernel(const float* __restrict__ A, int64_t lda, int64_t R, int64_t Nc,
                     float* __restrict__ part) {
  // block (chunk, column-tile): 8 row-lanes x 32 columns
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 32 + cx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kColsumRows;
  const int64_t r1 = min(R, r0 + kColsumRows);
  float acc = 0.f;
  if (c < Nc) {
#pragma unroll 8
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += A[r * lda + c];
  }
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < Nc) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][cx];
    part[static_cast<int64_t>(blockIdx.y) * Nc + c] = t;
  }
}

// out[r,:] = sum_{p in [rowptr[r],rowptr[r+1])} X[idx[p],:].  Block (r, s) sums chunk s of row r:
// 256 threads = H/4 float4 lanes x 1024/H entry lanes, combined in a fixed order; S > 1 chunks per
// row go through `partial` and a second fixed-order pass (few long rows: embedding gradients).
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ X, const int32_t* __restrict__ rowptr,
                   const int32_t* __restrict__ idx, int H, int S, float* __restrict__ out) {
  __shared__ float4 s_acc[256];
  const int vec = H / 4, EL = 256 / vec;
  const int l = threadIdx.x % vec, el = threadIdx.x / vec;
  const int64_t r = blockIdx.x;
  const int s = blockIdx.y;
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  const int32_t per = (end - beg + S - 1) / S;
  const int32_t p0 = min(beg + s * per, end), p1 = min(p0 + per, end);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int32_t p = p0 + el; p < p1; p += EL) {
    const float4 v = *reinterpret_cast<const float4*>(X + static_cast<int64_t>(idx[p]) * H + 4 * l);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  if (el == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < EL; ++k) {
      const float4 v = s_acc[k * vec + l];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (r * S + s) * H + 4 * l) = t;
  }
}
// out[r, c] = sum_s partial[(r*S+s)*H + c], s ascending
__global__ void segment_sum_final_kernel(const float* __restrict__ partial, int64_t RH, int H, int S,
                                         float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= RH) return;
  const int64_t r = i / H, c = i % H;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(r * S + s) * H + c];
  out[i] = a;
}
static int segment_split(int64_t entries, int64_t R) {
  const int64_t avg = R > 0 ? cdiv(entries, R) : 0;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(64, cdiv(avg, 512))));
}

static int pick_splits(int64_t R, int64_t tiles) {
  // enough K-slices to fill the machine, each at least 64 rows deep; depends on
  // sizes only, so the reduction order is reproducible.
  int64_t want = std::max<int64_t>(1, (4 * kNumSMs) / std::max<int64_t>(tiles, 1));
  int64_t cap = std::max<int64_t>(1, R / 64);
  return static_cast<int>(std::min<int64_t>(std::min(want, cap), 1024));
}

}  // namespace qot

using namespace qot;

extern "C" int qot_gemm(const float* A, int64_t a_rs, int64_t a_cs, const int64_t* gather,
                        const float* B, int64_t b_rs, int64_t b_cs, const float* bias,
                        float* C, int64_t ldc, int64_t M, int64_t Nc, int64_t K, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(M >= 0 && Nc >= 0 && K >= 0, "qot_gemm: negative size");
  if (M == 0 || Nc == 0) return QOT_OK;
  QOT_REQUIRE(A && B && C, "qot_gemm: null operand");
  QOT_REQUIRE(ldc >= Nc, "qot_gemm: ldc < Nc");
  dim3 grid(static_cast<unsigned>(cdiv(M, BM)), static_cast<unsigned>(cdiv(Nc, BN)), 1);
  QOT_REQUIRE(grid.y <= 65535u, "qot_gemm: Nc too large for one launch (%lld columns)", (long long)Nc);
  gemm_kernel<<<grid, 256, 0, stream>>>(A, a_rs, a_cs, gather, B, b_rs, b_cs, bias, C, ldc, M, Nc, K, std::max<int64_t>(K, 1));
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_wgrad_workspace_bytes(int64_t R, int64_t Mo, int64_t No) {
  if (R < 0 || Mo < 0 || No < 0) return 0;
  const int splits = pick_splits(R, cdiv(Mo, BM) * cdiv(No, BN));
  return align_up(static_cast<size_t>(splits) * Mo * No * 4) + 256;
}

extern "C" int qot_wgrad(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t R,
                         int64_t Mo, int64_t No, float* C, int64_t ldc,
                         void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(R >= 0 && Mo >= 0 && No >= 0, "qot_wgrad: negative size");
  if (Mo == 0 || No == 0) return QOT_OK;
  QOT_REQUIRE(C && (R == 0 || (A && B)), "qot_wgrad: null operand");
  QOT_REQUIRE(ws && ws_bytes >= qot_wgrad_workspace_bytes(R, Mo, No), "qot_wgrad: workspace too small");
  const int64_t tiles = cdiv(Mo, BM) * cdiv(No, BN);
  const int splits = pick_splits(R, tiles);
  const int64_t kps = cdiv(std::max<int64_t>(R, 1), splits);
  float* part = static_cast<float*>(ws);
  dim3 grid(static_cast<unsigned>(cdiv(Mo, BM)), static_cast<unsigned>(cdiv(No, BN)), static_cast<unsigned>(splits));
  if (splits == 1) {
    gemm_kernel<<<grid, 256, 0, stream>>>(A, 1, lda, nullptr, B, ldb, 1, nullptr, C, ldc, Mo, No, R, kps);
    QOT_LAUNCH_CHECK();
    return QOT_OK;
  }
  gemm_kernel<<<grid, 256, 0, stream>>>(A, 1, lda, nullptr, B, ldb, 1, nullptr, part, No, Mo, No, R, kps);
  QOT_LAUNCH_CHECK();
  const int64_t n = Mo * No;
  reduce_partials_kernel<<<static_cast<unsigned>(cdiv(n, 8)), 256, 0, stream>>>(part, n, splits, C, ldc, No);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_colsum_workspace_bytes(int64_t R, int64_t Nc) {
  if (R < 0 || Nc < 0) return 0;
  return align_up(static_cast<size_t>(cdiv(std::max<int64_t>(R, 1), kColsumRows)) * Nc * 4) + 256;
}

extern "C" int qot_colsum(const float* A, int64_t lda, int64_t R, int64_t Nc, float* out,
                          void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(R >= 0 && Nc >= 0, "qot_colsum: negative size");
  if (Nc == 0) return QOT_OK;
  QOT_REQUIRE(out && (R == 0 || A), "qot_colsum: null operand");
  QOT_REQUIRE(ws && ws_bytes >= qot_colsum_workspace_bytes(R, Nc), "qot_colsum: workspace too small");
  const int64_t chunks = cdiv(std::max<int64_t>(R, 1), kColsumRows);
  QOT_REQUIRE(chunks <= 65535, "qot_colsum: too many rows for one launch");
  float* part = static_cast<float*>(ws);
  dim3 grid(static_cast<unsigned>(cdiv(Nc, 32)), static_cast<unsigned>(chunks));
  colsum_stage1_kernel<<<grid, 256, 0, stream>>>(A, lda, R, Nc, part);
  QOT_LAUNCH_CHECK();
  reduce_partials_kernel<<<static_cast<unsigned>(cdiv(Nc, 8)), 256, 0, stream>>>(
      part, Nc, static_cast<int>(chunks), out, Nc, Nc);
  QOT_LAUNCH_CHECK();
  return QOT_OK;
}

extern "C" size_t qot_segment_sum_workspace_bytes(int64_t entries, int64_t R, int64_t H) {
  if (entries < 0 || R < 0 || H <= 0) return 0;
  return align_up(static_cast<size_t>(std::max<int64_t>(R, 1)) * segment_split(entries, R) * H * 4) + 256;
}

extern "C" int qot_segment_sum(const float* X, const int32_t* rowptr, const int32_t* idx, int64_t entries,
                               int64_t R, int64_t H, float* out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QOT_REQUIRE(R >= 0 && entries >= 0 && (H == 16 || H == 32 || H == 64 || H == 128 || H == 256),
              "qot_segment_sum: H must be 16/32/64/128/256");
  if (R == 0) return QOT_OK;
  QOT_REQUIRE(rowptr && out && (entries == 0 || (X && idx)), "qot_segment_sum: null operand");
  QOT_REQUIRE(ws && ws_bytes >= qot_segment_sum_workspace_bytes(entries, R, H), "qot_segment_sum: workspace too small");
  QOT_REQUIRE(R <= 0x7fffffffll, "qot_segment_sum: too many rows for one launch");
  const int S = segment_split(entries, R);
  float* dst = S == 1 ? out : static_cast<float*>(ws);
  segment_sum_kernel<<<dim3(static_cast<unsigned>(R), S), 256, 0, stream>>>(X, rowptr, idx, static_cast<int>(H), S, dst);
  QOT_LAUNCH_CHECK();
  if (S > 1) {
    segment_sum_final_kernel<<<static_cast<unsigned>(cdiv(R * H, 256)), 256, 0, stream>>>(dst, R * H, static_cast<int>(H), S, out);
    QOT_LAUNCH_CHECK();
  }
  return QOT_OK;
}
