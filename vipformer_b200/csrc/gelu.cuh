// gelu.cuh -- exact-erf GELU and its derivative (shared by the GEMM epilogue and the streaming kernels)
#pragma once

namespace vpf {

// exact-erf GELU (nn.GELU default, partseg.py:196) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below
// the bf16 rounding of the stored result); one MUFU.EX2 + one MUFU.RCP instead of the ~30-instruction erff().
// MUFU.RCP / MUFU.EX2 as single instructions: __frcp_rn() is a CALL with an IEEE fix-up path and __expf() carries
// denormal scaling -- together they made the streaming GELU kernels issue-bound at 40 SASS instructions per element.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void erf_parts(float x, float &erf_v, float &gauss) {
  const float z = fabsf(x) * 0.70710678118654752f;   // erf(x / sqrt 2), exp(-x^2 / 2)
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  gauss = ex2_approx(z * z * -1.4426950408889634f);   // exp(-z^2) = 2^(-z^2 log2 e)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  erf_v = copysignf(fmaf(-p * t, gauss, 1.0f), x);
}
__device__ __forceinline__ float gelu_f(float x) {
  float er, ga;
  erf_parts(x, er, ga);
  return 0.5f * x * (1.0f + er);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float er, ga;
  erf_parts(x, er, ga);
  return 0.5f * (1.0f + er) + x * 0.39894228040143268f * ga;
}

}  // namespace vpf
