// Developer microbenchmark: issue cadence of the legacy mma.sync path on sm_100a for the shapes the
// lightpath readout head could use, and whether it starves a co-resident FFMA warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// mode 0 tf32 k8, 1 f16 k16, 2 bf16 k16; `ffma_warps` extra warps per block run an FFMA loop instead
template <int MODE>
__global__ void rate_kernel(int iters, int mma_warps, long long* cycles_mma, long long* cycles_ffma, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned a[4] = {0x3f800000u + lane, 0x3f810000u, 0x3f820000u, 0x3f830000u};
  float acc[8][4] = {};
  float f[8] = {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f};
  __syncthreads();
  const long long t0 = clock64();
  if (warp < mma_warps) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {                     // 8 independent accumulators: issue-rate, not latency
        if (MODE == 0) mma_tf32(acc[u], a, a[1], a[2]);
        if (MODE == 1) mma_f16(acc[u], a, a[1], a[2]);
        if (MODE == 2) mma_bf16(acc[u], a, a[1], a[2]);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) cycles_mma[blockIdx.x * 32 + warp] = t1 - t0;
  } else {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = fmaf(f[u], 1.0001f, 0.5f);
    }
    const long long t1 = clock64();
    if (lane == 0) cycles_ffma[blockIdx.x * 32 + warp] = t1 - t0;
  }
  float s = 0.f;
  for (int u = 0; u < 8; ++u) s += acc[u][0] + acc[u][3] + f[u];
  if (s == 123.456f) sink[0] = s;
}

template <int MODE>
void run(const char* name, int mma_warps, int ffma_warps) {
  long long *cm, *cf;
  float* sink;
  cudaMalloc(&cm, 148 * 32 * 8); cudaMalloc(&cf, 148 * 32 * 8); cudaMalloc(&sink, 4);
  cudaMemset(cm, 0, 148 * 32 * 8); cudaMemset(cf, 0, 148 * 32 * 8);
  const int iters = 2000, threads = 32 * (mma_warps + ffma_warps);
  rate_kernel<MODE><<<148, threads>>>(iters, mma_warps, cm, cf, sink);
  rate_kernel<MODE><<<148, threads>>>(iters, mma_warps, cm, cf, sink);
  cudaDeviceSynchronize();
  long long hm[32], hf[32];
  cudaMemcpy(hm, cm, sizeof(hm), cudaMemcpyDeviceToHost);
  cudaMemcpy(hf, cf, sizeof(hf), cudaMemcpyDeviceToHost);
  double mm = 0, ff = 0;
  for (int w = 0; w < mma_warps; ++w) mm = hm[w] > mm ? hm[w] : mm;
  for (int w = mma_warps; w < mma_warps + ffma_warps; ++w) ff = hf[w] > ff ? hf[w] : ff;
  const double per_smsp = (mma_warps + 3) / 4;            // warps sharing one scheduler
  printf("%-5s mma_warps=%2d ffma_warps=%2d : %.1f cycles per MMA per warp, %.1f cycles per MMA per scheduler", name, mma_warps,
         ffma_warps, mm / (iters * 8.0), mm / (iters * 8.0) / per_smsp);
  if (ffma_warps) printf(" | FFMA warp: %.2f cycles per FFMA", ff / (iters * 32.0));
  printf("\n");
  cudaFree(cm); cudaFree(cf); cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) { run<0>("tf32", w, 0); run<1>("f16", w, 0); run<2>("bf16", w, 0); }
  run<0>("tf32", 0, 4);                                   // FFMA alone: one warp per scheduler
  run<0>("tf32", 4, 4);                                   // one MMA warp + one FFMA warp per scheduler
  run<1>("f16", 4, 4);
  run<0>("tf32", 8, 8);
  return 0;
}
