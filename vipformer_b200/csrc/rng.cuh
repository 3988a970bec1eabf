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

// ---- Residual / Dropout masks (partseg.py:208-213, dp1 of partseg.py:401): byte-granular.  One 32-bit hash serves the four
// consecutive elements 4g .. 4g+3 of the flattened [rows, cols] tensor, element e keeps iff byte (e & 3) of
// hash(key, e >> 2) >= thr8, thr8 = round(256 p): p = 0.5 is exact, p = 0.1 becomes 26/256 = 0.1016 (the same
// quantisation as the attention-probability masks), and the survivors are scaled by 256 / (256 - thr8), the exact
// inverse of the keep probability.  A quarter of the integer work of a hash per element, which is what the streaming
// dropout-gradient kernel and the residual GEMM epilogue were spending their issue slots on.  Restated in oracle/rng.py.
__host__ __device__ __forceinline__ uint32_t threshold8(float p) {
  if (!(p > 0.f)) return 0u;
  const uint32_t t = (uint32_t)(p * 256.f + 0.5f);
  return t > 255u ? 255u : t;
}
__host__ __device__ __forceinline__ float scale8(uint32_t thr8) { return 256.f / (float)(256u - thr8); }
__host__ __device__ __forceinline__ uint32_t hash4(uint32_t key, uint32_t group) { return mix32(group * 0x9e3779b1u ^ key); }
__host__ __device__ __forceinline__ bool keep8(uint32_t key, uint32_t e, uint32_t thr8) {
  return ((hash4(key, e >> 2) >> ((e & 3u) * 8u)) & 0xffu) >= thr8;
}
// v[0..NV) = elements ebase .. ebase+NV-1 of one row; `aligned`: ebase % 4 == 0 (one hash per four values)
template <int NV>
__device__ __forceinline__ void drop_values(float (&v)[NV], uint32_t key, uint32_t ebase, uint32_t thr8, float scale, bool aligned) {
  static_assert(NV % 4 == 0, "NV must be a multiple of 4");
  if (aligned) {
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) {
      const uint32_t h = hash4(key, (ebase >> 2) + k);
#pragma unroll
      for (int b = 0; b < 4; ++b) v[4 * k + b] = ((h >> (8 * b)) & 0xffu) >= thr8 ? v[4 * k + b] * scale : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = keep8(key, ebase + j, thr8) ? v[j] * scale : 0.f;
  }
}


// ---- DropPath (timm.models.layers.DropPath as used by Residual, partseg.py:201-213): one keep decision per SAMPLE,
// survivors scaled by 1 / (1 - p).  32-bit threshold: drop iff hash(key, b) < p * 2^32.  Restated in oracle/rng.py.
__host__ __device__ __forceinline__ uint32_t threshold32(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool keep_sample(uint32_t key, uint32_t b, uint32_t thr32) {
  return mix32(b * 0x9e3779b1u ^ key) >= thr32;
}

}  // namespace rng
}  // namespace vpf
