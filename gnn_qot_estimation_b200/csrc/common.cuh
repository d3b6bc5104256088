// Shared host/device helpers for libqot_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qot_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libqot_b200 targets sm_100a (B200) only"
#endif

namespace qot {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);   // api.cu (thread-local message)

#define QOT_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::qot::set_error(__VA_ARGS__);              \
      return QOT_E_BADARG;                        \
    }                                             \
  } while (0)

#define QOT_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ::qot::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                       cudaGetErrorString(e_));                                     \
      return QOT_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define QOT_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t e_ = cudaPeekAtLastError();                                         \
    if (e_ != cudaSuccess) {                                                        \
      ::qot::set_error("%s:%d launch -> %s", __FILE__, __LINE__,                    \
                       cudaGetErrorString(e_));                                     \
      return QOT_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr unsigned kFull = 0xffffffffu;

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute is a per-DEVICE setting: `done` keeps one bit per device ordinal, so a process
// driving several GPUs opts every one of them in (and two threads racing only repeat an idempotent call).
template <typename F>
static inline int once_per_device(std::atomic<unsigned long long>& done, F&& set) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return QOT_E_CUDA;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return QOT_OK;
  const int rc = set();
  if (rc == QOT_OK) done.fetch_or(bit, std::memory_order_release);
  return rc;
}

// Carves aligned sub-buffers out of the caller's workspace.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T));
    return r;
  }
};

// Device-wide exclusive scan of n int32 values (out has n+1 entries, out[n] = total).
// `add` is added to every input element before scanning (self-loop reservation).
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t add, int32_t* out, int64_t n, void* ws,
                       cudaStream_t stream);

#ifdef __CUDACC__
// ---- device helpers --------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

#endif

}  // namespace qot
