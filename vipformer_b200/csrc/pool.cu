// pool.cu -- max/mean pooling, K=3 linears (too thin for tensor cores), image patchify.
// Reference semantics:
//   torch.max over the 32 neighbours of a group      utils.py:180,188   (first max wins)
//   cat(x.max(1)[0], x.mean(1))                      partseg.py:547
//   nn.Linear(3, 64|128) / Conv1d(3, 64, 1)          classifier.py:32  partseg.py:499  utils.py:154
//   Rearrange 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)'   partseg.py:632
#include "common.cuh"
#include "rng.cuh"

namespace vpf {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ---------------------------------------------------------------- group max
// x bf16 [G,S,C] -> max over S (first index wins); out_bf16 / out_f32 optional; argmax u8 [G,C]
__global__ void __launch_bounds__(128)
group_max_fwd_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ out_bf16, float *__restrict__ out_f32,
                     uint8_t *__restrict__ argmax, int S, int C) {
  const int g = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const bf16 *p = x + (size_t)g * S * C + c;
  float best = __bfloat162float(p[0]);
  int bi = 0;
  for (int s = 1; s < S; ++s) {
    const float v = __bfloat162float(p[(size_t)s * C]);
    if (v > best) { best = v; bi = s; }
  }
  const size_t o = (size_t)g * C + c;
  if (out_bf16) out_bf16[o] = __float2bfloat16(best);
  if (out_f32) out_f32[o] = best;
  argmax[o] = (uint8_t)bi;
}

// dx[g,s,c] (+)= (s == argmax[g,c]) ? dout[g,c] : 0
template <typename Tdo>
__global__ void __launch_bounds__(128)
group_max_bwd_kernel(const Tdo *__restrict__ dout, const uint8_t *__restrict__ argmax, bf16 *__restrict__ dx,
                     int accumulate, int S, int C) {
  const int g = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const size_t o = (size_t)g * C + c;
  const int am = argmax[o];
  float d;
  if constexpr (sizeof(Tdo) == 4) d = dout[o]; else d = __bfloat162float(dout[o]);
  bf16 *p = dx + (size_t)g * S * C + c;
  if (accumulate) {
    p[(size_t)am * C] = __float2bfloat16(__bfloat162float(p[(size_t)am * C]) + d);
  } else {
    const bf16 z = __float2bfloat16(0.f), dv = __float2bfloat16(d);
    for (int s = 0; s < S; ++s) p[(size_t)s * C] = (s == am) ? dv : z;
  }
}

// sum over the S rows of each group: x bf16 [G,S,C] -> out_bf16 [G,C] (optional) and out_f32 (optional)
__global__ void __launch_bounds__(128)
group_sum_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ out_bf16, float *__restrict__ out_f32, int S, int C) {
  const int g = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const bf16 *p = x + (size_t)g * S * C + c;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(p[(size_t)s * C]);
  const size_t o = (size_t)g * C + c;
  if (out_bf16) out_bf16[o] = __float2bfloat16(acc);
  if (out_f32) out_f32[o] = acc;
}

// --------------------------------------------------------------- token pool
// x fp32 [B,L,D] -> out fp32 [B,2D] = (max_l || mean_l); argmax int32 [B,D]
__global__ void __launch_bounds__(128)
token_pool_fwd_kernel(const float *__restrict__ x, float *__restrict__ out, int *__restrict__ argmax, int L, int D) {
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= D) return;
  const float *p = x + (size_t)b * L * D + c;
  float best = p[0], sum = p[0];
  int bi = 0;
  for (int l = 1; l < L; ++l) {
    const float v = p[(size_t)l * D];
    sum += v;
    if (v > best) { best = v; bi = l; }
  }
  out[(size_t)b * 2 * D + c] = best;
  out[(size_t)b * 2 * D + D + c] = sum / (float)L;
  argmax[(size_t)b * D + c] = bi;
}
__global__ void __launch_bounds__(128)
token_pool_bwd_kernel(const float *__restrict__ dout, const int *__restrict__ argmax, float *__restrict__ dx, int L, int D) {
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= D) return;
  const float dmax = dout[(size_t)b * 2 * D + c], dmean = dout[(size_t)b * 2 * D + D + c] / (float)L;
  const int am = argmax[(size_t)b * D + c];
  float *p = dx + (size_t)b * L * D + c;
  for (int l = 0; l < L; ++l) p[(size_t)l * D] = dmean + (l == am ? dmax : 0.f);
}

// ------------------------------------------------------------ K = 3 linears
// y = ((w.p + b) * scale + shift) ; pre (optional, bf16) gets y; act (optional, bf16) gets act(y)
// one thread = 8 consecutive output channels of one row (16-byte stores); Co % 8 == 0
__global__ void __launch_bounds__(256)
linear3_fwd_kernel(const float *__restrict__ p, int ldp, const float *__restrict__ w, const float *__restrict__ b,
                   const float *__restrict__ scale, const float *__restrict__ shift, bf16 *__restrict__ pre,
                   bf16 *__restrict__ actout, int act, long long R, int Co) {
  const size_t total8 = (size_t)R * Co / 8;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total8; i += (size_t)gridDim.x * 256) {
    const size_t e = i * 8;
    const int c0 = (int)(e % Co);
    const size_t r = e / Co;
    const float *q = p + r * ldp;
    const float x0 = q[0], x1 = q[1], x2 = q[2];
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      float t = fmaf(w[c * 3 + 2], x2, fmaf(w[c * 3 + 1], x1, w[c * 3] * x0)) + b[c];
      if (scale) t = t * scale[c] + shift[c];
      y[k] = t;
    }
    uint4 u;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
    if (pre) {
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(y[2 * k], y[2 * k + 1]);
      *reinterpret_cast<uint4 *>(pre + e) = u;
    }
    if (actout) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (act == VPF_ACT_RELU) y[k] = fmaxf(y[k], 0.f);
        else if (act == VPF_ACT_GELU) y[k] = gelu_exact(y[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(y[2 * k], y[2 * k + 1]);
      *reinterpret_cast<uint4 *>(actout + e) = u;
    }
  }
}

// column statistics of y = w.p + b without materialising it: stats[0..Co) += sum, stats[Co..2Co) += sumsq
__global__ void __launch_bounds__(256)
linear3_stats_kernel(const float *__restrict__ p, int ldp, const float *__restrict__ w, const float *__restrict__ b,
                     double *__restrict__ stats, long long R, int Co, int rows_per_cta) {
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  const int lx = min(Co, 256), ny = 256 / lx, ty = threadIdx.x / lx;   // ty strides rows: no idle threads when Co < 256
  if (ty >= ny) return;
  for (int c = threadIdx.x % lx; c < Co; c += lx) {
    const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2], bb = b[c];
    double a = 0.0, q = 0.0;
    for (long long r0 = row0 + ty; r0 < row1; r0 += 64LL * ny) {
      float pa = 0.f, pq = 0.f;
      const long long r1 = min(row1, r0 + 64LL * ny);
      for (long long r = r0; r < r1; r += ny) {
        const float *x = p + (size_t)r * ldp;
        const float y = fmaf(w2, x[2], fmaf(w1, x[1], w0 * x[0])) + bb;
        pa += y; pq += y * y;
      }
      a += pa; q += pq;
    }
    atomicAdd(stats + c, a);
    atomicAdd(stats + Co + c, q);
  }
}

// dW[c][j] += sum_r dy[r,c] p[r,j];  db[c] += sum_r dy[r,c]
template <typename Tdy>
__global__ void __launch_bounds__(256)
linear3_bwd_kernel(const Tdy *__restrict__ dy, const float *__restrict__ p, int ldp, float *__restrict__ dW,
                   float *__restrict__ db, long long R, int Co, int rows_per_cta) {
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  const int lx = min(Co, 256), ny = 256 / lx, ty = threadIdx.x / lx;
  if (ty >= ny) return;
  for (int c = threadIdx.x % lx; c < Co; c += lx) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
    for (long long r = row0 + ty; r < row1; r += ny) {
      float d;
      if constexpr (sizeof(Tdy) == 4) d = dy[(size_t)r * Co + c]; else d = __bfloat162float(dy[(size_t)r * Co + c]);
      const float *x = p + (size_t)r * ldp;
      a0 += d * x[0]; a1 += d * x[1]; a2 += d * x[2]; ab += d;
    }
    atomicAdd(dW + c * 3 + 0, a0); atomicAdd(dW + c * 3 + 1, a1); atomicAdd(dW + c * 3 + 2, a2);
    atomicAdd(db + c, ab);
  }
}

// BN(train)+ReLU backward through y = w.p + b recomputed on the fly (Group2Emb first_conv.0-2, utils.py:154-156).
// phase 1: red[0..Co) += sum dyb, red[Co..2Co) += sum dyb*xhat
// phase 2: dy1 = scale*(dyb - m1 - xhat*m2): dW += dy1^T p, db += sum dy1   (dy1 never stored)
template <int PHASE>
__global__ void __launch_bounds__(256)
linear3_bn_bwd_kernel(const bf16 *__restrict__ dh, const float *__restrict__ p, int ldp, const float *__restrict__ w,
                      const float *__restrict__ b, const float *__restrict__ scale, const float *__restrict__ shift,
                      const float *__restrict__ mean, const float *__restrict__ rstd, double *__restrict__ red,
                      float *__restrict__ dW, float *__restrict__ db, long long R, int Co, int rows_per_cta) {
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  const int lx = min(Co, 256), ny = 256 / lx, ty = threadIdx.x / lx;
  if (ty >= ny) return;
  for (int c = threadIdx.x % lx; c < Co; c += lx) {
    const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2], bb = b[c];
    const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
    float m1 = 0.f, m2 = 0.f;
    if (PHASE == 2) { m1 = (float)(red[c] / (double)R); m2 = (float)(red[Co + c] / (double)R); }
    double a = 0.0, q = 0.0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
    for (long long r0 = row0 + ty; r0 < row1; r0 += 64LL * ny) {
      float pa = 0.f, pq = 0.f;
      const long long r1 = min(row1, r0 + 64LL * ny);
      for (long long r = r0; r < r1; r += ny) {
        const float *x = p + (size_t)r * ldp;
        const float y = fmaf(w2, x[2], fmaf(w1, x[1], w0 * x[0])) + bb;
        float d = __bfloat162float(dh[(size_t)r * Co + c]);
        if (!(y * sc + sh > 0.f)) d = 0.f;
        const float xh = (y - mu) * rs;
        if (PHASE == 1) { pa += d; pq += d * xh; }
        else {
          const float g = sc * (d - m1 - xh * m2);
          a0 += g * x[0]; a1 += g * x[1]; a2 += g * x[2]; ab += g;
        }
      }
      a += pa; q += pq;
    }
    if (PHASE == 1) { atomicAdd(red + c, a); atomicAdd(red + Co + c, q); }
    else {
      atomicAdd(dW + c * 3 + 0, a0); atomicAdd(dW + c * 3 + 1, a1); atomicAdd(dW + c * 3 + 2, a2);
      atomicAdd(db + c, ab);
    }
  }
}
// ------------------------------------------------------------------------------------------------------------------
// Channel-stationary versions of the K=3 kernels.  A thread owns CH consecutive output channels for the whole kernel
// (weights, bias and BatchNorm constants in registers, 16- / 8-byte accesses to the [R, Co] bf16 tensor) and strides
// the rows; Co/CH lanes cover one row, 256/(Co/CH) rows are in flight per pass, U passes are unrolled.  The scalar
// kernels above (one channel per thread, every constant re-read from global memory per element, 2-byte accesses) ran
// at 9-15 % of the HBM roofline; they stay as the fallback for channel counts the vector layout does not divide.
struct L3Map { int lanes, ny, tx, ty, c0; };
template <int CH>
__device__ __forceinline__ L3Map l3_map(int Co) {
  L3Map m;
  m.lanes = Co / CH;
  m.ny = 256 / m.lanes;
  m.tx = threadIdx.x % m.lanes;
  m.ty = threadIdx.x / m.lanes;
  m.c0 = m.tx * CH;
  return m;
}
static inline bool l3_vec_ok(int Co, int CH) { return Co % CH == 0 && Co / CH <= 256 && 256 % (Co / CH) == 0; }

__device__ __forceinline__ void unpack8_bf16(const uint4 u, float *f) {
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8_bf16(const float *f) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  return u;
}

__global__ void __launch_bounds__(256)
linear3_fwd_vec_kernel(const float *__restrict__ p, int ldp, const float *__restrict__ w, const float *__restrict__ b,
                       const float *__restrict__ scale, const float *__restrict__ shift, bf16 *__restrict__ pre,
                       bf16 *__restrict__ actout, int act, long long R, int Co, int rows_per_cta) {
  constexpr int CH = 8, U = 4;
  const L3Map m = l3_map<CH>(Co);
  float w0[CH], w1[CH], w2[CH], bb[CH], sc[CH], sh[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = m.c0 + k;
    w0[k] = w[c * 3]; w1[k] = w[c * 3 + 1]; w2[k] = w[c * 3 + 2]; bb[k] = b[c];
    sc[k] = scale ? scale[c] : 1.f; sh[k] = scale ? shift[c] : 0.f;
  }
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long r = row0 + m.ty; r < row1; r += (long long)U * m.ny) {
    float x0[U], x1[U], x2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * m.ny;
      if (rr < row1) { const float *q = p + (size_t)rr * ldp; x0[u] = q[0]; x1[u] = q[1]; x2[u] = q[2]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * m.ny;
      if (rr >= row1) continue;
      float y[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        float t = fmaf(w2[k], x2[u], fmaf(w1[k], x1[u], w0[k] * x0[u])) + bb[k];   // same order as the scalar kernel
        if (scale) t = t * sc[k] + sh[k];
        y[k] = t;
      }
      const size_t e = (size_t)rr * Co + m.c0;
      if (pre) *reinterpret_cast<uint4 *>(pre + e) = pack8_bf16(y);
      if (actout) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (act == VPF_ACT_RELU) y[k] = fmaxf(y[k], 0.f);
          else if (act == VPF_ACT_GELU) y[k] = gelu_exact(y[k]);
        }
        *reinterpret_cast<uint4 *>(actout + e) = pack8_bf16(y);
      }
    }
  }
}

// fixed-order combine of the ny row-lanes of a CTA through shared memory, then one atomic per column and CTA
template <typename T, int NV>
__device__ __forceinline__ void l3_cta_reduce(const L3Map &m, int Co, int CH, const T (*vals)[8], T *const *outs, T *smem) {
  // vals[v][k]: value v of this thread's channel k; smem holds [ny][Co] of T
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
    for (int k = 0; k < CH; ++k) smem[m.ty * Co + m.c0 + k] = vals[v][k];
    __syncthreads();
    for (int c = threadIdx.x; c < Co; c += 256) {
      T t = 0;
      for (int y = 0; y < m.ny; ++y) t += smem[y * Co + c];
      atomicAdd(outs[v] + c, t);
    }
  }
}

__global__ void __launch_bounds__(256)
linear3_stats_vec_kernel(const float *__restrict__ p, int ldp, const float *__restrict__ w, const float *__restrict__ b,
                         double *__restrict__ stats, long long R, int Co, int rows_per_cta) {
  constexpr int CH = 8, U = 4;
  extern __shared__ double s_l3d[];   // [ny][Co]
  const L3Map m = l3_map<CH>(Co);
  float w0[CH], w1[CH], w2[CH], bb[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) { const int c = m.c0 + k; w0[k] = w[c * 3]; w1[k] = w[c * 3 + 1]; w2[k] = w[c * 3 + 2]; bb[k] = b[c]; }
  double acc[2][8];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[0][k] = acc[1][k] = 0.0;
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long rb = row0 + m.ty; rb < row1; rb += 16LL * U * m.ny) {   // fp32 partials over <= 64 rows, then fp64
    float pa[CH], pq[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) pa[k] = pq[k] = 0.f;
    const long long re = min(row1, rb + 16LL * U * m.ny);
    for (long long r = rb; r < re; r += (long long)U * m.ny) {
      float x0[U], x1[U], x2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * m.ny;
        if (rr < re) { const float *q = p + (size_t)rr * ldp; x0[u] = q[0]; x1[u] = q[1]; x2[u] = q[2]; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (long long)u * m.ny >= re) continue;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const float y = fmaf(w2[k], x2[u], fmaf(w1[k], x1[u], w0[k] * x0[u])) + bb[k];
          pa[k] += y; pq[k] += y * y;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) { acc[0][k] += pa[k]; acc[1][k] += pq[k]; }
  }
  double *outs[2] = {stats, stats + Co};
  l3_cta_reduce<double, 2>(m, Co, CH, acc, outs, s_l3d);
}

// BN(train)+ReLU backward through the recomputed y = w.p + b (see linear3_bn_bwd_kernel): 4 channels per thread
template <int PHASE>
__global__ void __launch_bounds__(256)
linear3_bn_bwd_vec_kernel(const bf16 *__restrict__ dh, const float *__restrict__ p, int ldp, const float *__restrict__ w,
                          const float *__restrict__ b, const float *__restrict__ scale, const float *__restrict__ shift,
                          const float *__restrict__ mean, const float *__restrict__ rstd, double *__restrict__ red,
                          float *__restrict__ dW, float *__restrict__ db, long long R, int Co, int rows_per_cta) {
  constexpr int CH = 4, U = 4;
  extern __shared__ double s_l3d[];   // [ny][Co] doubles (phase 1) / floats (phase 2)
  const L3Map m = l3_map<CH>(Co);
  float w0[CH], w1[CH], w2[CH], bb[CH], sc[CH], sh[CH], rs[CH], nmr[CH], m1[CH], m2[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = m.c0 + k;
    w0[k] = w[c * 3]; w1[k] = w[c * 3 + 1]; w2[k] = w[c * 3 + 2]; bb[k] = b[c];
    sc[k] = scale[c]; sh[k] = shift[c]; rs[k] = rstd[c]; nmr[k] = mean[c];
    m1[k] = PHASE == 2 ? (float)(red[c] / (double)R) : 0.f;
    m2[k] = PHASE == 2 ? (float)(red[Co + c] / (double)R) : 0.f;
  }
  double dacc[2][8];
  float facc[4][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { dacc[0][k] = dacc[1][k] = 0.0; facc[0][k] = facc[1][k] = facc[2][k] = facc[3][k] = 0.f; }
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long rb = row0 + m.ty; rb < row1; rb += 16LL * U * m.ny) {
    float pa[CH], pq[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) pa[k] = pq[k] = 0.f;
    const long long re = min(row1, rb + 16LL * U * m.ny);
    for (long long r = rb; r < re; r += (long long)U * m.ny) {
      float x0[U], x1[U], x2[U];
      uint2 dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * m.ny;
        if (rr < re) {
          const float *q = p + (size_t)rr * ldp;
          x0[u] = q[0]; x1[u] = q[1]; x2[u] = q[2];
          dv[u] = __ldg(reinterpret_cast<const uint2 *>(dh + (size_t)rr * Co + m.c0));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (long long)u * m.ny >= re) continue;
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&dv[u]);
        const float2 d01 = __bfloat1622float2(h[0]), d23 = __bfloat1622float2(h[1]);
        const float dd[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const float y = fmaf(w2[k], x2[u], fmaf(w1[k], x1[u], w0[k] * x0[u])) + bb[k];
          float d = dd[k];
          if (!(y * sc[k] + sh[k] > 0.f)) d = 0.f;
          const float xh = (y - nmr[k]) * rs[k];
          if (PHASE == 1) { pa[k] += d; pq[k] += d * xh; }
          else {
            const float g = sc[k] * (d - m1[k] - xh * m2[k]);
            facc[0][k] += g * x0[u]; facc[1][k] += g * x1[u]; facc[2][k] += g * x2[u]; facc[3][k] += g;
          }
        }
      }
    }
    if (PHASE == 1) {
#pragma unroll
      for (int k = 0; k < CH; ++k) { dacc[0][k] += pa[k]; dacc[1][k] += pq[k]; }
    }
  }
  if (PHASE == 1) {
    double *outs[2] = {red, red + Co};
    l3_cta_reduce<double, 2>(m, Co, CH, dacc, outs, s_l3d);
  } else {
    // dW is [Co][3]: combine per column in shared memory, then strided atomics
    float *sf = reinterpret_cast<float *>(s_l3d);
    for (int v = 0; v < 4; ++v) {
      __syncthreads();
      for (int k = 0; k < CH; ++k) sf[m.ty * Co + m.c0 + k] = facc[v][k];
      __syncthreads();
      for (int c = threadIdx.x; c < Co; c += 256) {
        float t = 0.f;
        for (int y = 0; y < m.ny; ++y) t += sf[y * Co + c];
        if (v < 3) atomicAdd(dW + c * 3 + v, t); else atomicAdd(db + c, t);
      }
    }
  }
}

// dW[c][j] += sum_r dy[r,c] p[r,j];  db[c] += sum_r dy[r,c]   (dy bf16, 8 channels per thread)
__global__ void __launch_bounds__(256)
linear3_bwd_vec_kernel(const bf16 *__restrict__ dy, const float *__restrict__ p, int ldp, float *__restrict__ dW,
                       float *__restrict__ db, long long R, int Co, int rows_per_cta) {
  constexpr int CH = 8, U = 4;
  extern __shared__ double s_l3d[];
  const L3Map m = l3_map<CH>(Co);
  float facc[4][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) facc[0][k] = facc[1][k] = facc[2][k] = facc[3][k] = 0.f;
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long r = row0 + m.ty; r < row1; r += (long long)U * m.ny) {
    float x0[U], x1[U], x2[U];
    uint4 dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * m.ny;
      if (rr < row1) {
        const float *q = p + (size_t)rr * ldp;
        x0[u] = q[0]; x1[u] = q[1]; x2[u] = q[2];
        dv[u] = __ldg(reinterpret_cast<const uint4 *>(dy + (size_t)rr * Co + m.c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + (long long)u * m.ny >= row1) continue;
      float d[8];
      unpack8_bf16(dv[u], d);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        facc[0][k] += d[k] * x0[u]; facc[1][k] += d[k] * x1[u]; facc[2][k] += d[k] * x2[u]; facc[3][k] += d[k];
      }
    }
  }
  float *sf = reinterpret_cast<float *>(s_l3d);
  for (int v = 0; v < 4; ++v) {
    __syncthreads();
    for (int k = 0; k < CH; ++k) sf[m.ty * Co + m.c0 + k] = facc[v][k];
    __syncthreads();
    for (int c = threadIdx.x; c < Co; c += 256) {
      float t = 0.f;
      for (int y = 0; y < m.ny; ++y) t += sf[y * Co + c];
      if (v < 3) atomicAdd(dW + c * 3 + v, t); else atomicAdd(db + c, t);
    }
  }
}

__global__ void bn_param_grad_kernel(const double *__restrict__ red, float *__restrict__ dgamma, float *__restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dgamma[c] += (float)red[C + c]; dbeta[c] += (float)red[c]; }
}

// ----------------------------------------------------------------- patchify
// img fp32 [B,H,W,Ci] (NHWC, what CrossFormer_img_mp.forward receives) or [B,Ci,H,W] (NCHW, what the data loader
// yields before pretrain.py:179 permutes it) -> out bf16 [B*(H/P)*(W/P), P*P*Ci] in (p1 p2 c) order
__global__ void __launch_bounds__(256)
patchify_kernel(const float *__restrict__ img, bf16 *__restrict__ out, int H, int W, int Ci, int P, int nchw, size_t total) {
  const int nw = W / P, nh = H / P, pd = P * P * Ci;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int k = (int)(e % pd);
    const size_t row = e / pd;
    const int pw = (int)(row % nw), ph = (int)((row / nw) % nh);
    const size_t b = row / ((size_t)nw * nh);
    const int c = k % Ci, p2 = (k / Ci) % P, p1 = k / (Ci * P);
    const size_t y = (size_t)ph * P + p1, x = (size_t)pw * P + p2;
    const size_t src = nchw ? ((b * Ci + c) * H + y) * W + x : ((b * H + y) * W + x) * Ci + c;
    out[e] = __float2bfloat16(img[src]);
  }
}

// NCHW source: walk the INPUT in memory order (x fastest) so the 4-byte reads coalesce; the 2-byte writes scatter with a
// 6-byte stride inside a patch row and are merged by L2 (the whole bf16 output is 32 MB).  The output-ordered kernel
// above reads an NCHW image with a stride of H*W floats between consecutive threads (148 us vs ~15 us of traffic).
__global__ void __launch_bounds__(256)
patchify_nchw_kernel(const float *__restrict__ img, bf16 *__restrict__ out, int H, int W, int Ci, int P, size_t total) {
  const int nw = W / P, nh = H / P, pd = P * P * Ci;
  for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (size_t)gridDim.x * 256) {
    const int x = (int)(t % W);
    const size_t t1 = t / W;
    const int y = (int)(t1 % H);
    const size_t t2 = t1 / H;
    const int c = (int)(t2 % Ci);
    const size_t b = t2 / Ci;
    const size_t row = (b * nh + y / P) * nw + x / P;
    const int k = ((y % P) * P + (x % P)) * Ci + c;
    out[row * pd + k] = __float2bfloat16(__ldg(img + t));
  }
}

// out = alpha * (a + b)   /   batch-sum reduce: out[r % rows] += x[r]
__global__ void add_scale_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, float alpha, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = alpha * (a[i] + (b ? b[i] : 0.f));
}

// dst[r, c] = alpha * src[r, c] over a [rows, cols] window of two row-strided fp32 arrays
__global__ void copy2d_kernel(const float *__restrict__ src, int lds, float *__restrict__ dst, int ldd, long long rows, int cols, float alpha) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t r = i / cols, c = i - r * cols;
    dst[r * ldd + c] = alpha * src[r * lds + c];
  }
}

// out[i] = x[i] * s[0]  (s on the device: an upstream scalar gradient)
__global__ void scale_by_kernel(const float *__restrict__ x, const float *__restrict__ s, float *__restrict__ out, size_t n) {
  const float f = s[0];
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = x[i] * f;
}

// eval-mode BatchNorm folded into the preceding 1x1 convolution (inference): Wout[n, k] = bf16(scale[n] * W[n, k]),
// bout[n] = scale[n] * b[n] + shift[n]
__global__ void bn_fold_kernel(const float *__restrict__ W, const float *__restrict__ b, const float *__restrict__ scale,
                               const float *__restrict__ shift, __nv_bfloat16 *__restrict__ Wout, float *__restrict__ bout, int N, int K) {
  const size_t n = (size_t)N * K;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int r = (int)(i / K);
    Wout[i] = __float2bfloat16(W[i] * scale[r]);
    if (i - (size_t)r * K == 0) bout[r] = (b ? b[r] : 0.f) * scale[r] + shift[r];
  }
}

static inline int grid_for(size_t total) { return (int)min((size_t)num_sms() * 8, ceil_div(total, (size_t)256)); }
static inline int rows_per_cta_for(long long R) { return (int)max((long long)64, ceil_div(R, (long long)num_sms() * 8)); }

// out[n] += sum_k v[k] * W[k, n]   (fp32 vector times a bf16 [K, ldw] matrix window; K, N a few hundred)
// Group2Emb backward uses it for a bias gradient: colsum(dY . W) = colsum(dY) . W, which replaces a pass over a 0.5 GB tensor.
__global__ void __launch_bounds__(128)
vecmat_bf16_kernel(const float *__restrict__ v, const bf16 *__restrict__ W, int ldw, int K, int N, float *__restrict__ out) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(v[k], __bfloat162float(W[(size_t)k * ldw + n]), acc);
  out[n] += acc;
}

// DropPath scales of one Residual: s[b] = keep(b) ? 1 / (1 - p) : 0   (one decision per sample, rng.cuh keep_sample)
__global__ void droppath_scales_kernel(const unsigned long long *__restrict__ seed_ptr, uint32_t op_id, float p, int B,
                                       float *__restrict__ s) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  const uint32_t key = rng::make_key(seed_ptr ? *seed_ptr : 0ull, op_id);
  s[b] = rng::keep_sample(key, (uint32_t)b, rng::threshold32(p)) ? 1.f / (1.f - p) : 0.f;
}
// out[r, :] = x[r, :] * s[r / L]  (fp32 rows of D % 4 == 0 channels; out may alias x): DropPath on a [B*L, D] token matrix
__global__ void __launch_bounds__(256)
row_scale_kernel(const float *__restrict__ x, const float *__restrict__ s, int L, float *__restrict__ out, long long T, int D4) {
  const size_t n = (size_t)T * D4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float f = s[(i / D4) / L];
    float4 v = reinterpret_cast<const float4 *>(x)[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    reinterpret_cast<float4 *>(out)[i] = v;
  }
}

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_group_max_fwd(const void *x_bf16, void *out_bf16, float *out_f32, uint8_t *argmax, int G, int S, int C, void *stream) {
  VPF_REQUIRE(x_bf16 && argmax && (out_bf16 || out_f32), "group_max_fwd: null pointer");
  VPF_REQUIRE(S >= 1 && S <= 255, "group_max_fwd: S=%d unsupported", S);
  if (G == 0 || C == 0) return VPF_OK;
  group_max_fwd_kernel<<<dim3(G, ceil_div(C, 128)), 128, 0, (cudaStream_t)stream>>>((const bf16 *)x_bf16, (bf16 *)out_bf16, out_f32, argmax, S, C);
  return check_launch("group_max_fwd_kernel");
}

int vpf_group_max_bwd(const void *dout, int dout_bf16, const uint8_t *argmax, void *dx_bf16, int accumulate, int G, int S, int C, void *stream) {
  VPF_REQUIRE(dout && argmax && dx_bf16, "group_max_bwd: null pointer");
  if (G == 0 || C == 0) return VPF_OK;
  dim3 grid(G, ceil_div(C, 128));
  if (dout_bf16) group_max_bwd_kernel<bf16><<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16 *)dout, argmax, (bf16 *)dx_bf16, accumulate, S, C);
  else group_max_bwd_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float *)dout, argmax, (bf16 *)dx_bf16, accumulate, S, C);
  return check_launch("group_max_bwd_kernel");
}

int vpf_group_sum(const void *x_bf16, void *out_bf16, float *out_f32, int G, int S, int C, void *stream) {
  VPF_REQUIRE(x_bf16 && (out_bf16 || out_f32), "group_sum: null pointer");
  if (G == 0 || C == 0) return VPF_OK;
  group_sum_kernel<<<dim3(G, ceil_div(C, 128)), 128, 0, (cudaStream_t)stream>>>((const bf16 *)x_bf16, (bf16 *)out_bf16, out_f32, S, C);
  return check_launch("group_sum_kernel");
}

int vpf_vecmat_bf16(const float *v, const void *W_bf16, int ldw, int K, int N, float *out, void *stream) {
  VPF_REQUIRE(v && W_bf16 && out && ldw >= N, "vecmat_bf16: bad arguments");
  if (K == 0 || N == 0) return VPF_OK;
  vecmat_bf16_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(v, (const bf16 *)W_bf16, ldw, K, N, out);
  return check_launch("vecmat_bf16_kernel");
}

int vpf_droppath_scales(const unsigned long long *seed_ptr, unsigned int op_id, float p, int B, float *scales, void *stream) {
  VPF_REQUIRE(scales && p >= 0.f && p < 1.f, "droppath_scales: bad arguments (p=%f)", p);
  if (B == 0) return VPF_OK;
  droppath_scales_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(seed_ptr, op_id, p, B, scales);
  return check_launch("droppath_scales_kernel");
}

int vpf_row_scale(const float *x, const float *scales, int L, float *out, long long T, int D, void *stream) {
  VPF_REQUIRE(x && scales && out && L >= 1 && D % 4 == 0, "row_scale: bad arguments (L=%d D=%d)", L, D);
  VPF_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "row_scale: pointers must be 16-byte aligned");
  if (T == 0 || D == 0) return VPF_OK;
  row_scale_kernel<<<grid_for((size_t)T * (D / 4)), 256, 0, (cudaStream_t)stream>>>(x, scales, L, out, T, D / 4);
  return check_launch("row_scale_kernel");
}

int vpf_token_pool_fwd(const float *x, float *out, int *argmax, int B, int L, int D, void *stream) {
  VPF_REQUIRE(x && out && argmax, "token_pool_fwd: null pointer");
  if (B == 0) return VPF_OK;
  token_pool_fwd_kernel<<<dim3(B, ceil_div(D, 128)), 128, 0, (cudaStream_t)stream>>>(x, out, argmax, L, D);
  return check_launch("token_pool_fwd_kernel");
}

int vpf_token_pool_bwd(const float *dout, const int *argmax, float *dx, int B, int L, int D, void *stream) {
  VPF_REQUIRE(dout && dx && argmax, "token_pool_bwd: null pointer");
  if (B == 0) return VPF_OK;
  token_pool_bwd_kernel<<<dim3(B, ceil_div(D, 128)), 128, 0, (cudaStream_t)stream>>>(dout, argmax, dx, L, D);
  return check_launch("token_pool_bwd_kernel");
}

int vpf_linear3_fwd(const float *p, int ldp, const float *w, const float *b, const float *scale, const float *shift,
                    void *pre_bf16, void *act_bf16, int act, long long R, int Co, void *stream) {
  VPF_REQUIRE(p && w && b && (pre_bf16 || act_bf16), "linear3_fwd: null pointer");
  VPF_REQUIRE((scale == nullptr) == (shift == nullptr), "linear3_fwd: scale/shift must come together");
  VPF_REQUIRE(Co % 8 == 0, "linear3_fwd: Co=%d must be a multiple of 8", Co);
  if (R == 0) return VPF_OK;
  if (l3_vec_ok(Co, 8)) {
    const int ny = 256 / (Co / 8);
    const int rpc = (int)max((long long)ny * 4, ceil_div(R, (long long)num_sms() * 8));
    linear3_fwd_vec_kernel<<<(unsigned)ceil_div(R, (long long)rpc), 256, 0, (cudaStream_t)stream>>>(p, ldp, w, b, scale, shift, (bf16 *)pre_bf16, (bf16 *)act_bf16, act, R, Co, rpc);
    return check_launch("linear3_fwd_vec_kernel");
  }
  linear3_fwd_kernel<<<grid_for((size_t)R * Co / 8), 256, 0, (cudaStream_t)stream>>>(p, ldp, w, b, scale, shift, (bf16 *)pre_bf16, (bf16 *)act_bf16, act, R, Co);
  return check_launch("linear3_fwd_kernel");
}

int vpf_linear3_stats(const float *p, int ldp, const float *w, const float *b, double *stats, long long R, int Co, void *stream) {
  VPF_REQUIRE(p && w && b && stats, "linear3_stats: null pointer");
  if (R == 0) return VPF_OK;
  const int rpc = rows_per_cta_for(R);
  if (l3_vec_ok(Co, 8)) {
    const int smem = (256 / (Co / 8)) * Co * (int)sizeof(double);
    VPF_REQUIRE(smem <= 48 * 1024, "linear3_stats: Co=%d too wide", Co);
    linear3_stats_vec_kernel<<<(unsigned)ceil_div(R, (long long)rpc), 256, smem, (cudaStream_t)stream>>>(p, ldp, w, b, stats, R, Co, rpc);
    return check_launch("linear3_stats_vec_kernel");
  }
  linear3_stats_kernel<<<(unsigned)ceil_div(R, (long long)rpc), 256, 0, (cudaStream_t)stream>>>(p, ldp, w, b, stats, R, Co, rpc);
  return check_launch("linear3_stats_kernel");
}

int vpf_linear3_bwd(const void *dy, int dy_bf16, const float *p, int ldp, float *dW, float *db, long long R, int Co, void *stream) {
  VPF_REQUIRE(dy && p && dW && db, "linear3_bwd: null pointer");
  if (R == 0) return VPF_OK;
  const int rpc = rows_per_cta_for(R);
  const unsigned grid = (unsigned)ceil_div(R, (long long)rpc);
  if (dy_bf16 && l3_vec_ok(Co, 8) && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const int smem = (256 / (Co / 8)) * Co * (int)sizeof(float);
    linear3_bwd_vec_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16 *)dy, p, ldp, dW, db, R, Co, rpc);
    return check_launch("linear3_bwd_vec_kernel");
  }
  if (dy_bf16) linear3_bwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16 *)dy, p, ldp, dW, db, R, Co, rpc);
  else linear3_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)dy, p, ldp, dW, db, R, Co, rpc);
  return check_launch("linear3_bwd_kernel");
}

int vpf_linear3_bn_bwd(const void *dh_bf16, const float *p, int ldp, const float *w, const float *b, const float *scale,
                       const float *shift, const float *mean, const float *rstd, double *red, float *dW, float *db,
                       float *dgamma, float *dbeta, long long R, int Co, void *stream) {
  VPF_REQUIRE(dh_bf16 && p && w && b && scale && shift && mean && rstd && red && dW && db && dgamma && dbeta, "linear3_bn_bwd: null pointer");
  if (R == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VPF_CUDA_TRY(cudaMemsetAsync(red, 0, sizeof(double) * 2 * Co, st));
  const int rpc = rows_per_cta_for(R);
  const unsigned grid = (unsigned)ceil_div(R, (long long)rpc);
  if (l3_vec_ok(Co, 4) && (reinterpret_cast<uintptr_t>(dh_bf16) & 7) == 0 && (256 / (Co / 4)) * Co * 8 <= 48 * 1024) {
    const int smem = (256 / (Co / 4)) * Co * (int)sizeof(double);
    linear3_bn_bwd_vec_kernel<1><<<grid, 256, smem, st>>>((const bf16 *)dh_bf16, p, ldp, w, b, scale, shift, mean, rstd, red, dW, db, R, Co, rpc);
    VPF_TRY(check_launch("linear3_bn_bwd_vec_kernel<1>"));
    linear3_bn_bwd_vec_kernel<2><<<grid, 256, smem, st>>>((const bf16 *)dh_bf16, p, ldp, w, b, scale, shift, mean, rstd, red, dW, db, R, Co, rpc);
    VPF_TRY(check_launch("linear3_bn_bwd_vec_kernel<2>"));
  } else {
    linear3_bn_bwd_kernel<1><<<grid, 256, 0, st>>>((const bf16 *)dh_bf16, p, ldp, w, b, scale, shift, mean, rstd, red, dW, db, R, Co, rpc);
    VPF_TRY(check_launch("linear3_bn_bwd_kernel<1>"));
    linear3_bn_bwd_kernel<2><<<grid, 256, 0, st>>>((const bf16 *)dh_bf16, p, ldp, w, b, scale, shift, mean, rstd, red, dW, db, R, Co, rpc);
    VPF_TRY(check_launch("linear3_bn_bwd_kernel<2>"));
  }
  bn_param_grad_kernel<<<ceil_div(Co, 128), 128, 0, st>>>(red, dgamma, dbeta, Co);
  return check_launch("bn_param_grad_kernel");
}

int vpf_patchify(const float *img, void *out_bf16, int B, int H, int W, int Ci, int P, int nchw, void *stream) {
  VPF_REQUIRE(img && out_bf16, "patchify: null pointer");
  VPF_REQUIRE(P > 0 && H % P == 0 && W % P == 0, "patchify: image %dx%d not divisible by patch %d", H, W, P);
  const size_t total = (size_t)B * H * W * Ci;
  if (total == 0) return VPF_OK;
  if (nchw) {
    patchify_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, (bf16 *)out_bf16, H, W, Ci, P, total);
    return check_launch("patchify_nchw_kernel");
  }
  patchify_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, (bf16 *)out_bf16, H, W, Ci, P, nchw, total);
  return check_launch("patchify_kernel");
}

int vpf_bn_fold(const float *W, const float *b, const float *scale, const float *shift, void *Wout_bf16, float *bout, int N,
                int K, void *stream) {
  VPF_REQUIRE(W && scale && shift && Wout_bf16 && bout, "bn_fold: null pointer");
  if (N == 0 || K == 0) return VPF_OK;
  bn_fold_kernel<<<grid_for((size_t)N * K), 256, 0, (cudaStream_t)stream>>>(W, b, scale, shift, (__nv_bfloat16 *)Wout_bf16, bout, N, K);
  return check_launch("bn_fold_kernel");
}

int vpf_scale_by(const float *x, const float *s, float *out, long long n, void *stream) {
  VPF_REQUIRE(x && s && out, "scale_by: null pointer");
  if (n == 0) return VPF_OK;
  scale_by_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, s, out, (size_t)n);
  return check_launch("scale_by_kernel");
}

int vpf_copy2d(const float *src, int lds, float *dst, int ldd, long long rows, int cols, float alpha, void *stream) {
  VPF_REQUIRE(src && dst && lds >= cols && ldd >= cols, "copy2d: bad arguments");
  if (rows == 0 || cols == 0) return VPF_OK;
  copy2d_kernel<<<grid_for((size_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, cols, alpha);
  return check_launch("copy2d_kernel");
}

int vpf_add_scale(const float *a, const float *b, float *out, float alpha, long long n, void *stream) {
  VPF_REQUIRE(a && out, "add_scale: null pointer");
  if (n == 0) return VPF_OK;
  add_scale_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(a, b, out, alpha, (size_t)n);
  return check_launch("add_scale_kernel");
}

}  // extern "C"
