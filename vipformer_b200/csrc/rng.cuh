// rng.cuh -- counter-based dropout RNG shared by forward epilogues and backward kernels.
// keep(e) is a pure function of (seed, op_id, element index): the backward pass regenerates the
// forward mask instead of storing it.  Hash = lowbias32 finaliser over (index ^ key).
#pragma once
#include <stdint.h>

namespace vpf {
namespace rng {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t make_key(unsigned long long seed, uint32_t op_id) {
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9e3779b9u * (op_id + 1u)));
}
// P(drop) = p  <=>  hash < threshold
__host__ __device__ __forceinline__ uint32_t threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool keep(uint32_t key, uint32_t index, uint32_t thr) {
  return mix32(index * 0x9e3779b1u ^ key) >= thr;
}

}  // namespace rng
}  // namespace vpf
