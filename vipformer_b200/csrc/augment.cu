// augment.cu -- device-side point-cloud augmentation: the trans_1 / trans_2 chain of the reference's input pipeline
// (datasets/data.py:16-36, applied per cloud at data.py:109-112 by 18 CPU DataLoader workers) as ONE kernel in front of
// the tokenizer: normalise to the unit sphere (data_utils.py:206-221) -> scale (:56-66) -> rotate about y (:69-98) ->
// translate by a fraction of the bounding box (:156-171) -> clipped Gaussian jitter (:141-153) -> overwrite a random
// subset with point 0 (:179-193: the exact duplicates that make the tokenizer's tie-break rules matter).
// The random draws are INPUTS (per cloud: scale, angle, 3 translation fractions, drop ratio; per point: 3 jitter values
// and one uniform), drawn by the caller with a device generator, so the kernel is a pure function that the oracle
// (oracle/augment.py, pinned to the reference classes) can check on identical draws.
// One CTA per cloud; the cloud lives in shared memory; three block reductions (centroid, radius, bounding box).
#include "common.cuh"
#include "rng.cuh"

namespace vpf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// params [B, 6] = (scale, angle, tx, ty, tz, drop_ratio); jitter [B, N, 3] (N(0, std), before the clamp); drop_u [B, N]
__global__ void __launch_bounds__(256)
augment_kernel(const float *__restrict__ pts, const float *__restrict__ params, const float *__restrict__ jitter,
               const float *__restrict__ drop_u, float *__restrict__ out, int N, float clip) {
  extern __shared__ float sp[];          // [N][3]
  __shared__ float red[8][6];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float *p = pts + (size_t)b * N * 3;
  // ---- centroid
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = tid; i < N; i += 256) {
    const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    sp[3 * i] = x; sp[3 * i + 1] = y; sp[3 * i + 2] = z;
    sx += x; sy += y; sz += z;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { red[warp][0] = sx; red[warp][1] = sy; red[warp][2] = sz; }
  __syncthreads();
  float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { cx += red[w][0]; cy += red[w][1]; cz += red[w][2]; }
  const float invn = 1.f / (float)N;
  cx *= invn; cy *= invn; cz *= invn;
  // ---- radius
  float m2 = 0.f;
  for (int i = tid; i < N; i += 256) {
    const float x = sp[3 * i] - cx, y = sp[3 * i + 1] - cy, z = sp[3 * i + 2] - cz;
    sp[3 * i] = x; sp[3 * i + 1] = y; sp[3 * i + 2] = z;
    m2 = fmaxf(m2, x * x + y * y + z * z);
  }
  m2 = warp_max(m2);
  __syncthreads();                      // everyone is done with red[][0..2]
  if (lane == 0) red[warp][3] = m2;
  __syncthreads();
  float mm = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) mm = fmaxf(mm, red[w][3]);
  const float *pr = params + (size_t)b * 6;
  const float scl = pr[0], ang = pr[1], ratio = pr[5];
  const float m = sqrtf(mm);
  float sn, cs;
  sincosf(ang, &sn, &cs);
  // ---- normalise, scale, rotate about y (points @ R^T with R = [[c,0,s],[0,1,0],[-s,0,c]]); bounding box
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < N; i += 256) {
    const float x = sp[3 * i] / m * scl, y = sp[3 * i + 1] / m * scl, z = sp[3 * i + 2] / m * scl;
    const float xr = cs * x + sn * z, zr = -sn * x + cs * z;
    sp[3 * i] = xr; sp[3 * i + 1] = y; sp[3 * i + 2] = zr;
    lo[0] = fminf(lo[0], xr); hi[0] = fmaxf(hi[0], xr);
    lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
    lo[2] = fminf(lo[2], zr); hi[2] = fmaxf(hi[2], zr);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { lo[k] = -warp_max(-lo[k]); hi[k] = warp_max(hi[k]); }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; }
  }
  __syncthreads();
  float tr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float l = INFINITY, h = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) { l = fminf(l, red[w][k]); h = fmaxf(h, red[w][3 + k]); }
    tr[k] = pr[2 + k] * (h - l);
  }
  // ---- translate + jitter (point 0 first: the dropout overwrites with the FINAL point 0)
  const float *jt = jitter + (size_t)b * N * 3;
  for (int i = tid; i < N; i += 256) {
#pragma unroll
    for (int k = 0; k < 3; ++k) sp[3 * i + k] = sp[3 * i + k] + tr[k] + fminf(fmaxf(jt[3 * i + k], -clip), clip);
  }
  __syncthreads();
  const float p0x = sp[0], p0y = sp[1], p0z = sp[2];
  const float *du = drop_u + (size_t)b * N;
  float *o = out + (size_t)b * N * 3;
  for (int i = tid; i < N; i += 256) {
    const bool drop = du[i] <= ratio;
    o[3 * i] = drop ? p0x : sp[3 * i];
    o[3 * i + 1] = drop ? p0y : sp[3 * i + 1];
    o[3 * i + 2] = drop ? p0z : sp[3 * i + 2];
  }
}

// FPS start indices (utils.py:71): out[i] = floor(u * N), u = 32-bit hash of (seed, op_id, i) / 2^32
__global__ void draw_indices_kernel(const long long *__restrict__ state, uint32_t op_id, int n, int N, long long *__restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t key = rng::make_key((unsigned long long)state[1], op_id);
  const uint32_t h = rng::mix32((uint32_t)i * 0x9e3779b1u ^ key);
  out[i] = (long long)(((unsigned long long)h * (unsigned long long)N) >> 32);
}

}  // namespace vpf

using namespace vpf;

extern "C" int vpf_draw_indices(const long long *state, unsigned int op_id, int n, int N, long long *out, void *stream) {
  VPF_REQUIRE(state && out && N >= 1, "draw_indices: bad arguments (N=%d)", N);
  if (n == 0) return VPF_OK;
  draw_indices_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(state, op_id, n, N, out);
  return check_launch("draw_indices_kernel");
}

extern "C" size_t vpf_attention_bwd_workspace_bytes(int B, int H, int Lq) { return (size_t)B * H * Lq * sizeof(float); }
extern "C" size_t vpf_bn_bwd_workspace_bytes(int C) { return (size_t)3 * C * sizeof(double); }
extern "C" size_t vpf_linear3_bn_bwd_workspace_bytes(int Co) { return (size_t)2 * Co * sizeof(double); }
extern "C" size_t vpf_ntxent_logits_workspace_bytes(int n_r, int n_c) { return (size_t)n_r * n_c * sizeof(float); }
extern "C" size_t vpf_ntxent_grad_workspace_bytes(int n_r, int D) { return (size_t)n_r * D * sizeof(float); }

extern "C" int vpf_augment_clouds(const float *pts, const float *params, const float *jitter, const float *drop_u, float *out,
                                  int B, int N, float jitter_clip, void *stream) {
  VPF_REQUIRE(pts && params && jitter && drop_u && out, "augment_clouds: null pointer");
  VPF_REQUIRE(N >= 1 && N <= 16384, "augment_clouds: N=%d out of range (1..16384)", N);
  if (B == 0) return VPF_OK;
  const int smem = N * 3 * (int)sizeof(float);
  if (smem > 48 * 1024) VPF_CUDA_TRY(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  augment_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(pts, params, jitter, drop_u, out, N, jitter_clip);
  return check_launch("augment_kernel");
}
