// Developer microbenchmark: L2 -> SM read bandwidth of this GPU, streaming (coalesced 128-bit loads) and gathering
// (one 1 KB row per warp at a random offset, the access pattern of the cfg-5 edge kernels), on a buffer that fits L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bw l2_bw.cu && ./l2_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void stream_kernel(const float4* __restrict__ buf, size_t n4, int reps, float* sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = __ldcg(buf + i);                    // L2 only: the per-SM slice would otherwise sit in L1
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

// each warp reads `rows_per_warp` rows of 256 floats (1 KB) at pseudo-random row indices
__global__ void gather_kernel(const float4* __restrict__ buf, unsigned nrows, int rows_per_warp, float* sink) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  unsigned s = warp * 2654435761u + 12345u;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < rows_per_warp; r += 4) {
    float4 v[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s = s * 1664525u + 1013904223u;
      const float4* row = buf + static_cast<size_t>(s % nrows) * 64;
      v[j][0] = __ldcg(row + lane);
      v[j][1] = __ldcg(row + lane + 32);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc.x += v[j][0].x + v[j][1].x; acc.y += v[j][0].y + v[j][1].y; }
  }
  if (acc.x + acc.y == 123.456f) *sink = acc.x;
}

int main() {
  const size_t bytes = 64ull << 20;                       // 64 MB: inside the 126 MB L2
  float4* buf; float* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    stream_kernel<<<blocks, 512>>>(buf, bytes / 16, 2, sink);          // warm L2
    cudaEventRecord(e0);
    stream_kernel<<<blocks, 512>>>(buf, bytes / 16, 20, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("stream  %5d blocks x 512: %.0f GB/s\n", blocks, 20.0 * bytes / (ms * 1e6));
  }
  const unsigned nrows = bytes / 1024;
  for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
    const int rpw = 256;
    gather_kernel<<<blocks, 256>>>(buf, nrows, 16, sink);
    cudaEventRecord(e0);
    gather_kernel<<<blocks, 256>>>(buf, nrows, rpw, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("gather  %5d blocks x 256, 1 KB rows: %.0f GB/s\n", blocks, static_cast<double>(blocks) * 8 * rpw * 1024 / (ms * 1e6));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
