// partseg.cu -- the kernels part segmentation needs on top of the pre-training path:
//   * three_nn_kernel: PointNetFeaturePropagation's 3-nearest-centre search and inverse-distance weights
//     (vipformer/model/pointcloud/utils.py:223-229).  The reference SORTS the whole [B,N,S] distance matrix and keeps three
//     columns; here every point scans the S centres (shared memory) keeping the three smallest (distance, index) keys.
//   * interp3_fwd / interp3_bwd: interpolated[n, :] = sum_j w[n, j] * feats[idx[n, j], :]  (utils.py:230), written straight
//     into the bf16 GEMM operand of the propagation MLP with the point coordinates appended as extra K columns
//     (the reference's cat([points1, interpolated]), utils.py:233-234, reordered to keep the operand 16-byte aligned).
//   * bf16 token pooling (x.max(2), x.mean(2) over the groups, partseg.py:432-434), LeakyReLU(0.2) (partseg.py:393),
//     column permutation of the first propagation weight, strided bf16 copies.
#include "common.cuh"

namespace vpf {

typedef __nv_bfloat16 bf16;

// distance exactly as square_distance(xyz1, xyz2) states it (utils.py:138-140): (-2 * dot + |p|^2) + |c|^2
__device__ __forceinline__ float nn_dist(float px, float py, float pz, float p2, const float4 c) {
  const float dot = fmaf(pz, c.z, fmaf(py, c.y, __fmul_rn(px, c.x)));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), p2), c.w);
}

__global__ void __launch_bounds__(256)
three_nn_kernel(const float *__restrict__ pts, const float *__restrict__ ctr, int N, int S, int *__restrict__ idx,
                float *__restrict__ w) {
  extern __shared__ float4 s_c[];   // [S] (x, y, z, |c|^2)
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < S; i += 256) {
    const float *c = ctr + ((size_t)b * S + i) * 3;
    s_c[i] = make_float4(c[0], c[1], c[2], __fadd_rn(__fadd_rn(__fmul_rn(c[0], c[0]), __fmul_rn(c[1], c[1])), __fmul_rn(c[2], c[2])));
  }
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const float *p = pts + ((size_t)b * N + n) * 3;
  const float px = p[0], py = p[1], pz = p[2];
  const float p2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
  float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
  int i0 = 0, i1 = 0, i2 = 0;
  for (int j = 0; j < S; ++j) {          // ascending j + strict < keeps the lowest index among equal distances
    const float d = nn_dist(px, py, pz, p2, s_c[j]);
    if (d < d2) {
      if (d < d1) {
        d2 = d1; i2 = i1;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
        else { d1 = d; i1 = j; }
      } else { d2 = d; i2 = j; }
    }
  }
  if (S < 3) { if (S < 2) { d1 = d0; i1 = i0; } d2 = d1; i2 = i1; }
  const float r0 = 1.0f / (d0 + 1e-8f), r1 = 1.0f / (d1 + 1e-8f), r2 = 1.0f / (d2 + 1e-8f);   // utils.py:227-229
  const float nrm = r0 + r1 + r2;
  const size_t o = ((size_t)b * N + n) * 3;
  idx[o] = i0; idx[o + 1] = i1; idx[o + 2] = i2;
  w[o] = r0 / nrm; w[o + 1] = r1 / nrm; w[o + 2] = r2 / nrm;
}

// out[b*N + n, c] = sum_j w_j feats[b*S + idx_j, c] for c < C (8 channels per thread); columns C..C+2 = xyz, C+3..ld-1 = 0
__global__ void __launch_bounds__(256)
interp3_fwd_kernel(const bf16 *__restrict__ feats, int ldf, const int *__restrict__ idx, const float *__restrict__ w,
                   const float *__restrict__ pts, bf16 *__restrict__ out, int ldo, long long rows, int N, int S, int C) {
  const int chunks = ldo / 8;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= rows * chunks) return;
  const long long r = t / chunks;
  const int c0 = (int)(t - r * chunks) * 8;
  const int b = (int)(r / N);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  if (c0 < C) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float wj = w[r * 3 + j];
      const uint4 u = *reinterpret_cast<const uint4 *>(feats + ((size_t)b * S + idx[r * 3 + j]) * ldf + c0);
      const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        acc[2 * q] = fmaf(wj, f.x, acc[2 * q]);
        acc[2 * q + 1] = fmaf(wj, f.y, acc[2 * q + 1]);
      }
    }
  } else if (c0 == C) {
    acc[0] = pts[r * 3]; acc[1] = pts[r * 3 + 1]; acc[2] = pts[r * 3 + 2];
  }
  uint4 o;
  __nv_bfloat162 *ho = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) ho[q] = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
  *reinterpret_cast<uint4 *>(out + (size_t)r * ldo + c0) = o;
}

// dfeats[b*S + idx_j, c] += w_j * dout[r, c]   (fp32 atomics; dfeats row stride ldd)
__global__ void __launch_bounds__(256)
interp3_bwd_kernel(const bf16 *__restrict__ dout, int ldo, const int *__restrict__ idx, const float *__restrict__ w,
                   float *__restrict__ dfeats, int ldd, long long rows, int N, int S, int C) {
  const int chunks = C / 8;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= rows * chunks) return;
  const long long r = t / chunks;
  const int c0 = (int)(t - r * chunks) * 8;
  const int b = (int)(r / N);
  const uint4 u = *reinterpret_cast<const uint4 *>(dout + (size_t)r * ldo + c0);
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
  float g[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); g[2 * q] = f.x; g[2 * q + 1] = f.y; }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float wj = w[r * 3 + j];
    float *d = dfeats + ((size_t)b * S + idx[r * 3 + j]) * ldd + c0;
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(d + q, wj * g[q]);
  }
}

// x bf16 [B, L, ld] -> out fp32 [B, 2C] = (max_l || mean_l), argmax int32 [B, C]
__global__ void __launch_bounds__(128)
token_pool_bf16_fwd_kernel(const bf16 *__restrict__ x, int ld, float *__restrict__ out, int *__restrict__ argmax, int L, int C) {
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const bf16 *p = x + (size_t)b * L * ld + c;
  float best = __bfloat162float(p[0]), sum = best;
  int bi = 0;
  for (int l = 1; l < L; ++l) {
    const float v = __bfloat162float(p[(size_t)l * ld]);
    sum += v;
    if (v > best) { best = v; bi = l; }
  }
  out[(size_t)b * 2 * C + c] = best;
  out[(size_t)b * 2 * C + C + c] = sum / (float)L;
  argmax[(size_t)b * C + c] = bi;
}
// dx fp32 [B, L, ld] += dmean / L + (l == argmax) * dmax
__global__ void __launch_bounds__(128)
token_pool_accum_bwd_kernel(const float *__restrict__ dout, const int *__restrict__ argmax, float *__restrict__ dx, int ld, int L, int C) {
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const float dmax = dout[(size_t)b * 2 * C + c], dmean = dout[(size_t)b * 2 * C + C + c] / (float)L;
  const int am = argmax[(size_t)b * C + c];
  float *p = dx + (size_t)b * L * ld + c;
  for (int l = 0; l < L; ++l) p[(size_t)l * ld] += dmean + (l == am ? dmax : 0.f);
}

// LeakyReLU: y = x > 0 ? x : slope * x ; dx = dy * (x > 0 ? 1 : slope)
__global__ void leaky_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, float slope, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) { const float v = x[i]; y[i] = v > 0.f ? v : slope * v; }
}
__global__ void leaky_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dx, float slope, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dx[i] = dy[i] * (x[i] > 0.f ? 1.f : slope);
}

// Wout bf16 [Co, ldo]: columns [0, C) = W[:, 3 + c], [C, C + 3) = W[:, 0..2], rest 0   (W fp32 [Co, 3 + C])
__global__ void permute_w_kernel(const float *__restrict__ W, bf16 *__restrict__ Wout, int Co, int C, int ldo) {
  const size_t n = (size_t)Co * ldo;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int o = (int)(i / ldo), c = (int)(i - (size_t)o * ldo);
    float v = 0.f;
    if (c < C) v = W[(size_t)o * (C + 3) + 3 + c];
    else if (c < C + 3) v = W[(size_t)o * (C + 3) + (c - C)];
    Wout[i] = __float2bfloat16(v);
  }
}
// dW fp32 [Co, 3 + C] += the same permutation of dWp fp32 [Co, ldo]
__global__ void unpermute_dw_kernel(const float *__restrict__ dWp, float *__restrict__ dW, int Co, int C, int ldo) {
  const size_t n = (size_t)Co * (C + 3);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int o = (int)(i / (C + 3)), c = (int)(i - (size_t)o * (C + 3));
    dW[i] += dWp[(size_t)o * ldo + (c < 3 ? C + c : c - 3)];
  }
}
// dst bf16 [rows, ldd] window <- src bf16 [rows, lds] window
__global__ void copy2d_bf16_kernel(const bf16 *__restrict__ src, int lds, bf16 *__restrict__ dst, int ldd, long long rows, int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t r = i / cols, c = i - r * cols;
    dst[r * ldd + c] = src[r * lds + c];
  }
}

static inline int grid_n(size_t total) { return (int)min((size_t)num_sms() * 8, ceil_div(total, (size_t)256)); }

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_three_nn(const float *pts, const float *centers, int B, int N, int S, int *idx, float *w, void *stream) {
  VPF_REQUIRE(pts && centers && idx && w && S >= 1 && S <= 4096, "three_nn: bad arguments (S=%d)", S);
  if (B == 0 || N == 0) return VPF_OK;
  three_nn_kernel<<<dim3(ceil_div(N, 256), B), 256, S * sizeof(float4), (cudaStream_t)stream>>>(pts, centers, N, S, idx, w);
  return check_launch("three_nn_kernel");
}
int vpf_interp3_fwd(const void *feats_bf16, int ldf, const int *idx, const float *w, const float *pts, void *out_bf16, int ldo,
                    int B, int N, int S, int C, void *stream) {
  VPF_REQUIRE(feats_bf16 && idx && w && pts && out_bf16, "interp3_fwd: null pointer");
  VPF_REQUIRE(C % 8 == 0 && ldf % 8 == 0 && ldo % 8 == 0 && ldo >= C + 8, "interp3_fwd: C=%d ldf=%d ldo=%d must be multiples of 8, ldo >= C + 8", C, ldf, ldo);
  const long long rows = (long long)B * N;
  if (rows == 0) return VPF_OK;
  interp3_fwd_kernel<<<(unsigned)ceil_div(rows * (ldo / 8), 256LL), 256, 0, (cudaStream_t)stream>>>((const bf16 *)feats_bf16, ldf, idx, w, pts, (bf16 *)out_bf16, ldo, rows, N, S, C);
  return check_launch("interp3_fwd_kernel");
}
int vpf_interp3_bwd(const void *dout_bf16, int ldo, const int *idx, const float *w, float *dfeats, int ldd, int B, int N, int S,
                    int C, void *stream) {
  VPF_REQUIRE(dout_bf16 && idx && w && dfeats && C % 8 == 0 && ldo % 8 == 0, "interp3_bwd: bad arguments");
  const long long rows = (long long)B * N;
  if (rows == 0) return VPF_OK;
  interp3_bwd_kernel<<<(unsigned)ceil_div(rows * (C / 8), 256LL), 256, 0, (cudaStream_t)stream>>>((const bf16 *)dout_bf16, ldo, idx, w, dfeats, ldd, rows, N, S, C);
  return check_launch("interp3_bwd_kernel");
}
int vpf_token_pool_bf16_fwd(const void *x_bf16, int ld, float *out, int *argmax, int B, int L, int C, void *stream) {
  VPF_REQUIRE(x_bf16 && out && argmax, "token_pool_bf16_fwd: null pointer");
  if (B == 0) return VPF_OK;
  token_pool_bf16_fwd_kernel<<<dim3(B, ceil_div(C, 128)), 128, 0, (cudaStream_t)stream>>>((const bf16 *)x_bf16, ld, out, argmax, L, C);
  return check_launch("token_pool_bf16_fwd_kernel");
}
int vpf_token_pool_accum_bwd(const float *dout, const int *argmax, float *dx, int ld, int B, int L, int C, void *stream) {
  VPF_REQUIRE(dout && argmax && dx, "token_pool_accum_bwd: null pointer");
  if (B == 0) return VPF_OK;
  token_pool_accum_bwd_kernel<<<dim3(B, ceil_div(C, 128)), 128, 0, (cudaStream_t)stream>>>(dout, argmax, dx, ld, L, C);
  return check_launch("token_pool_accum_bwd_kernel");
}
int vpf_leaky_relu_fwd(const float *x, float *y, float slope, long long n, void *stream) {
  VPF_REQUIRE(x && y, "leaky_relu_fwd: null pointer");
  if (n == 0) return VPF_OK;
  leaky_fwd_kernel<<<grid_n((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, y, slope, (size_t)n);
  return check_launch("leaky_fwd_kernel");
}
int vpf_leaky_relu_bwd(const float *dy, const float *x, float *dx, float slope, long long n, void *stream) {
  VPF_REQUIRE(dy && x && dx, "leaky_relu_bwd: null pointer");
  if (n == 0) return VPF_OK;
  leaky_bwd_kernel<<<grid_n((size_t)n), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, slope, (size_t)n);
  return check_launch("leaky_bwd_kernel");
}
int vpf_permute_w(const float *W, void *Wout_bf16, int Co, int C, int ldo, void *stream) {
  VPF_REQUIRE(W && Wout_bf16 && ldo >= C + 3, "permute_w: bad arguments");
  permute_w_kernel<<<grid_n((size_t)Co * ldo), 256, 0, (cudaStream_t)stream>>>(W, (bf16 *)Wout_bf16, Co, C, ldo);
  return check_launch("permute_w_kernel");
}
int vpf_unpermute_dw(const float *dWp, float *dW, int Co, int C, int ldo, void *stream) {
  VPF_REQUIRE(dWp && dW && ldo >= C + 3, "unpermute_dw: bad arguments");
  unpermute_dw_kernel<<<grid_n((size_t)Co * (C + 3)), 256, 0, (cudaStream_t)stream>>>(dWp, dW, Co, C, ldo);
  return check_launch("unpermute_dw_kernel");
}
int vpf_copy2d_bf16(const void *src, int lds, void *dst, int ldd, long long rows, int cols, void *stream) {
  VPF_REQUIRE(src && dst && lds >= cols && ldd >= cols, "copy2d_bf16: bad arguments");
  if (rows == 0 || cols == 0) return VPF_OK;
  copy2d_bf16_kernel<<<grid_n((size_t)rows * cols), 256, 0, (cudaStream_t)stream>>>((const bf16 *)src, lds, (bf16 *)dst, ldd, rows, cols);
  return check_launch("copy2d_bf16_kernel");
}

}  // extern "C"
