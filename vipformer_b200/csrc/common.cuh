// common.cuh -- shared host/device helpers for libvpf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/vpf.h"

namespace vpf {

// ---- error plumbing: no exceptions cross the C ABI -------------------------
char *last_error_buf();                 // thread-local, 512 bytes
int fail(int code, const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

inline int check_launch(const char *what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(VPF_ECUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  }
  return VPF_OK;
}

#define VPF_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess)                                                        \
      return ::vpf::fail(VPF_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define VPF_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::vpf::fail(VPF_EINVAL, __VA_ARGS__);   \
  } while (0)

#define VPF_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != VPF_OK) return _rc; \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Resident CTAs of `kernel` on the whole device (SMs x occupancy): the grid of a streaming kernel that strides over its
// rows.  A grid a little larger than this runs a second, mostly empty wave -- at 591 CTAs over 444 slots the LayerNorm
// backward lost a third of its bandwidth to that tail.  One occupancy query per (kernel, block, smem) call site.
template <typename K>
inline int resident_ctas(K kernel, int threads, size_t dyn_smem = 0) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  return num_sms() * per_sm;
}
#define VPF_RESIDENT_CTAS(var, kernel, threads, smem)              \
  static int var##_cached = 0;                                      \
  if (var##_cached == 0) var##_cached = ::vpf::resident_ctas(kernel, threads, smem); \
  const int var = var##_cached

}  // namespace vpf
