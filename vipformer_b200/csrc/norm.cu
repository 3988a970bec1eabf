// norm.cu -- LayerNorm / BatchNorm / dropout-gradient / column-reduction kernels (fp32 math,
// bf16 or fp32 storage).  Reference semantics:
//   nn.LayerNorm (eps 1e-5, biased variance)       partseg.py:101-102,129,194  classifier.py:33
//   nn.BatchNorm1d train mode (eps 1e-5, momentum 0.1, biased var to normalise, unbiased var
//   into running_var)                              utils.py:155,162  partseg.py:520,523
//   nn.Dropout inside Residual                     partseg.py:201-213
#include "common.cuh"
#include "rng.cuh"
#include "gelu.cuh"

namespace vpf {

template <typename T> __device__ __forceinline__ float ldf(const T *p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float *p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p, size_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T *p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float *p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, size_t i, float v) { p[i] = __float2bfloat16(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 4-wide accessors: lane l owns elements [(l + 32 v) * 4, +4) of a row  (16-byte fp32 / 8-byte bf16 accesses)
template <typename T> __device__ __forceinline__ float4 ld4(const T *p, size_t i);
template <> __device__ __forceinline__ float4 ld4<float>(const float *p, size_t i) { return *reinterpret_cast<const float4 *>(p + i); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p, size_t i) {
  const uint2 u = *reinterpret_cast<const uint2 *>(p + i);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void st4(T *p, size_t i, float4 v);
template <> __device__ __forceinline__ void st4<float>(float *p, size_t i, float4 v) { *reinterpret_cast<float4 *>(p + i) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16 *p, size_t i, float4 v) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162 *>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162 *>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2 *>(p + i) = u;
}

// ------------------------------------------------------------- LayerNorm fwd
// one warp per row; NPER = D/32 values per lane held in registers.  VEC: lane owns float4 groups (D % 128 == 0).
template <typename Tin, int NPER, bool VEC>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const Tin *__restrict__ x, const float *__restrict__ add, int add_rows, float *__restrict__ xsum,
              const float *__restrict__ gamma, const float *__restrict__ beta, __nv_bfloat16 *__restrict__ y,
              float *__restrict__ mean_out, float *__restrict__ rstd_out, int T, float eps, int relu) {
  constexpr int D = NPER * 32;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  float v[NPER];
  float s = 0.f;
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < NPER / 4; ++q) {
      const int c = (lane + 32 * q) * 4;
      float4 t = ld4<Tin>(x, (size_t)row * D + c);
      if (add) {
        const float4 a = ld4<float>(add, (size_t)(row % add_rows) * D + c);
        t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      }
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      s += (t.x + t.y) + (t.z + t.w);
      if (xsum) st4<float>(xsum, (size_t)row * D + c, t);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int c = lane + 32 * i;
      float t = ldf<Tin>(x, (size_t)row * D + c);
      if (add) t += add[(size_t)(row % add_rows) * D + c];
      v[i] = t;
      s += t;
      if (xsum) xsum[(size_t)row * D + c] = t;
    }
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < NPER; ++i) { const float d = v[i] - mean; q2 += d * d; }
  const float rstd = rsqrtf(warp_sum(q2) * (1.f / D) + eps);
  if (lane == 0) { if (mean_out) mean_out[row] = mean; if (rstd_out) rstd_out[row] = rstd; }
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < NPER / 4; ++q) {
      const int c = (lane + 32 * q) * 4;
      const float4 g = ld4<float>(gamma, c), b = ld4<float>(beta, c);
      float4 o = make_float4((v[4 * q] - mean) * rstd * g.x + b.x, (v[4 * q + 1] - mean) * rstd * g.y + b.y,
                             (v[4 * q + 2] - mean) * rstd * g.z + b.z, (v[4 * q + 3] - mean) * rstd * g.w + b.w);
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      st4<__nv_bfloat16>(y, (size_t)row * D + c, o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int c = lane + 32 * i;
      float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (relu) o = fmaxf(o, 0.f);
      y[(size_t)row * D + c] = __float2bfloat16(o);
    }
  }
}

// Row-striding variant for narrow rows (D <= 128): the grid is the resident set and every warp walks the rows with the NEXT
// row's loads issued before the current row is reduced -- at D = 64 (the input adapter's LayerNorm over every point) the
// one-row-per-warp kernel above spends its time on CTA turnover: 156 -> 125 us per 1 M rows.  At D = 256 the prefetch
// registers halve the occupancy and it measured slower (256 vs 213 us), so wide rows keep the kernel above.
template <typename Tin, int NPER, bool VEC>
__device__ __forceinline__ void ln_load_row(const Tin *__restrict__ x, const float *__restrict__ add, int add_rows, int row,
                                            int lane, float (&v)[NPER]) {
  constexpr int D = NPER * 32;
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < NPER / 4; ++q) {
      const int c = (lane + 32 * q) * 4;
      float4 t = ld4<Tin>(x, (size_t)row * D + c);
      if (add) {
        const float4 a = ld4<float>(add, (size_t)(row % add_rows) * D + c);
        t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      }
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int c = lane + 32 * i;
      float t = ldf<Tin>(x, (size_t)row * D + c);
      if (add) t += add[(size_t)(row % add_rows) * D + c];
      v[i] = t;
    }
  }
}

template <typename Tin, int NPER, bool VEC>
__global__ void __launch_bounds__(256)
ln_fwd_rows_kernel(const Tin *__restrict__ x, const float *__restrict__ add, int add_rows, float *__restrict__ xsum,
              const float *__restrict__ gamma, const float *__restrict__ beta, __nv_bfloat16 *__restrict__ y,
              float *__restrict__ mean_out, float *__restrict__ rstd_out, int T, float eps, int relu) {
  constexpr int D = NPER * 32;
  const int lane = threadIdx.x & 31, nwarps = gridDim.x * 8;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= T) return;
  auto col = [&](int i) { return VEC ? (lane + 32 * (i >> 2)) * 4 + (i & 3) : lane + 32 * i; };
  float gm[NPER], bt[NPER];
#pragma unroll
  for (int i = 0; i < NPER; ++i) { gm[i] = gamma[col(i)]; bt[i] = beta[col(i)]; }
  float v[NPER], vn[NPER];
  ln_load_row<Tin, NPER, VEC>(x, add, add_rows, row, lane, v);
  for (; row < T; row += nwarps) {
    const int next = row + nwarps;
    if (next < T) ln_load_row<Tin, NPER, VEC>(x, add, add_rows, next, lane, vn);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.f / D);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) { const float d = v[i] - mean; q2 += d * d; }
    const float rstd = rsqrtf(warp_sum(q2) * (1.f / D) + eps);
    if (lane == 0) { if (mean_out) mean_out[row] = mean; if (rstd_out) rstd_out[row] = rstd; }
    float o[NPER];
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      o[i] = (v[i] - mean) * rstd * gm[i] + bt[i];
      if (relu) o[i] = fmaxf(o[i], 0.f);
    }
    if constexpr (VEC) {
#pragma unroll
      for (int q = 0; q < NPER / 4; ++q) {
        const size_t e = (size_t)row * D + (lane + 32 * q) * 4;
        if (xsum) st4<float>(xsum, e, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        st4<__nv_bfloat16>(y, e, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const size_t e = (size_t)row * D + lane + 32 * i;
        if (xsum) xsum[e] = v[i];
        y[e] = __float2bfloat16(o[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NPER; ++i) v[i] = vn[i];
  }
}

// ------------------------------------------------------------- LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = gamma * dy  (dy masked by y > 0 when relu)
// dx_out = dres + dx;  dpos[row % pos_rows] += dx_out;  dgamma += sum dy*xhat;  dbeta += sum dy
// EMIT: the kernel also writes g = bf16(dropout_mask(dx_out) * scale) and accumulates its column sums -- the masked copy
// of the residual-stream gradient that the NEXT block down the backward chain needs as a GEMM operand (Residual,
// partseg.py:208-213: d/d(branch output) = mask * dx, bias gradient = its column sum).  That used to be a separate
// pass over dx (dropout_grad); with one hash per four elements and this kernel at ~20 % issue utilisation it rides along.
struct LnEmit {
  __nv_bfloat16 *g;
  float *colsum;
  const unsigned long long *seed;
  uint32_t op_id;
  float p;
};
template <typename Tdy, typename Tx, typename Tdx, int NPER, bool VEC, bool EMIT = false>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const Tdy *__restrict__ dy, const Tx *__restrict__ x, const __nv_bfloat16 *__restrict__ y_relu,
              const float *__restrict__ mean_in, const float *__restrict__ rstd_in, const float *__restrict__ gamma,
              const float *__restrict__ dres, Tdx *__restrict__ dx, float *__restrict__ dgamma,
              float *__restrict__ dbeta, float *__restrict__ dpos, int pos_rows, int T, int rows_per_cta, LnEmit em) {
  constexpr int D = NPER * 32;
  static_assert(!EMIT || VEC, "the emitting variant exists for the vector layout only");
  __shared__ float s_dg[D], s_db[D];
  uint32_t em_thr = 0, em_key = 0;
  float em_scale = 1.f;
  float gs[EMIT ? NPER : 1];
  if constexpr (EMIT) {
    em_thr = rng::threshold8(em.p);
    em_key = em_thr ? rng::make_key(em.seed ? *em.seed : 0ull, em.op_id) : 0u;
    em_scale = rng::scale8(em_thr);
#pragma unroll
    for (int i = 0; i < NPER; ++i) gs[i] = 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dg[NPER], db[NPER], gm[NPER];
  // element i of this lane lives at column col(i)
  auto col = [&](int i) { return VEC ? (lane + 32 * (i >> 2)) * 4 + (i & 3) : lane + 32 * i; };
#pragma unroll
  for (int i = 0; i < NPER; ++i) { dg[i] = db[i] = 0.f; gm[i] = gamma[col(i)]; }
  // Row order.  Default: the CTA owns rows [row0, row1), warps interleaved.  With a broadcast position-embedding
  // gradient (dpos [pos_rows, D], pos_rows < T, partseg.py:289-296 adds pos to every layer's input) each warp owns ONE
  // token index and walks the batch, so the sum over the batch stays in registers and leaves as one atomic per
  // (warp, column) instead of one per (row, column): 16.8 M -> 1.2 M atomics per launch at 512 clouds.
  const bool tok_major = dpos != nullptr && pos_rows < T;
  float dp[NPER];
#pragma unroll
  for (int i = 0; i < NPER; ++i) dp[i] = 0.f;
  int row, row1, rstep, tok = 0;
  if (tok_major) {
    const int gw = blockIdx.x * 8 + warp, nchunk = (gridDim.x * 8) / pos_rows;   // host: gridDim.x * 8 >= pos_rows
    tok = gw % pos_rows;
    const int chunk = gw / pos_rows;
    row = chunk < nchunk ? chunk * pos_rows + tok : T;
    rstep = nchunk * pos_rows;
    row1 = T;
  } else {
    row = blockIdx.x * rows_per_cta + warp;
    row1 = min(T, blockIdx.x * rows_per_cta + rows_per_cta);
    rstep = 8;
  }
  for (; row < row1; row += rstep) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const size_t base = (size_t)row * D;
    float d[NPER], xh[NPER], rsd[NPER];
    if constexpr (VEC) {
#pragma unroll
      for (int q = 0; q < NPER / 4; ++q) {
        const int c = (lane + 32 * q) * 4;
        const float4 a = ld4<Tdy>(dy, base + c), b = ld4<Tx>(x, base + c);
        d[4 * q] = a.x; d[4 * q + 1] = a.y; d[4 * q + 2] = a.z; d[4 * q + 3] = a.w;
        xh[4 * q] = b.x; xh[4 * q + 1] = b.y; xh[4 * q + 2] = b.z; xh[4 * q + 3] = b.w;
        if (dres) {
          const float4 r = ld4<float>(dres, base + c);
          rsd[4 * q] = r.x; rsd[4 * q + 1] = r.y; rsd[4 * q + 2] = r.z; rsd[4 * q + 3] = r.w;
        }
        if (y_relu) {
          const float4 yy = ld4<__nv_bfloat16>(y_relu, base + c);
          if (!(yy.x > 0.f)) d[4 * q] = 0.f;
          if (!(yy.y > 0.f)) d[4 * q + 1] = 0.f;
          if (!(yy.z > 0.f)) d[4 * q + 2] = 0.f;
          if (!(yy.w > 0.f)) d[4 * q + 3] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int c = lane + 32 * i;
        d[i] = ldf<Tdy>(dy, base + c);
        if (y_relu && !(__bfloat162float(y_relu[base + c]) > 0.f)) d[i] = 0.f;
        xh[i] = ldf<Tx>(x, base + c);
        if (dres) rsd[i] = dres[base + c];
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      xh[i] = (xh[i] - mean) * rstd;
      dg[i] += d[i] * xh[i];
      db[i] += d[i];
      d[i] *= gm[i];
      s1 += d[i];
      s2 += d[i] * xh[i];
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      d[i] = rstd * (d[i] - s1 - xh[i] * s2);
      if (dres) d[i] += rsd[i];
    }
    if constexpr (VEC) {
#pragma unroll
      for (int q = 0; q < NPER / 4; ++q) {
        const int c = (lane + 32 * q) * 4;
        const float4 o = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
        st4<Tdx>(dx, base + c, o);
        if constexpr (EMIT) {
          float f[4] = {o.x, o.y, o.z, o.w};
          if (em_thr) rng::drop_values<4>(f, em_key, (uint32_t)(base + c), em_thr, em_scale, true);
          st4<__nv_bfloat16>(em.g, base + c, make_float4(f[0], f[1], f[2], f[3]));
          gs[4 * q] += f[0]; gs[4 * q + 1] += f[1]; gs[4 * q + 2] += f[2]; gs[4 * q + 3] += f[3];
        }
        if (dpos) {
          if (!tok_major) {
            float4 pp = ld4<float>(dpos, base + c);
            pp.x += o.x; pp.y += o.y; pp.z += o.z; pp.w += o.w;
            st4<float>(dpos, base + c, pp);
          } else {
            dp[4 * q] += o.x; dp[4 * q + 1] += o.y; dp[4 * q + 2] += o.z; dp[4 * q + 3] += o.w;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int c = lane + 32 * i;
        stf<Tdx>(dx, base + c, d[i]);
        if (dpos) {
          if (!tok_major) dpos[base + c] += d[i];
          else dp[i] += d[i];
        }
      }
    }
  }
  if (tok_major) {
#pragma unroll
    for (int i = 0; i < NPER; ++i) atomicAdd(dpos + (size_t)tok * D + col(i), dp[i]);
  }
  if (dgamma) {
    for (int c = threadIdx.x; c < D; c += 256) { s_dg[c] = 0.f; s_db[c] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPER; ++i) { atomicAdd(&s_dg[col(i)], dg[i]); atomicAdd(&s_db[col(i)], db[i]); }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
      atomicAdd(dgamma + c, s_dg[c]);
      atomicAdd(dbeta + c, s_db[c]);
    }
  }
  if constexpr (EMIT) {
    if (em.colsum) {
      __syncthreads();
      for (int c = threadIdx.x; c < D; c += 256) s_dg[c] = 0.f;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NPER; ++i) atomicAdd(&s_dg[col(i)], gs[i]);
      __syncthreads();
      for (int c = threadIdx.x; c < D; c += 256) atomicAdd(em.colsum + c, s_dg[c]);
    }
  }
}

// --------------------------------------------- residual-branch gradient prep
// out_bf16 = dropout_mask(g) (same (seed, op_id, index) stream as the forward epilogue); colsum += sum_rows(out)
__global__ void __launch_bounds__(256)
dropout_grad_kernel(const float *__restrict__ g, __nv_bfloat16 *__restrict__ out, float *__restrict__ colsum,
                    float p, const unsigned long long *__restrict__ seed_ptr, uint32_t op_id, int T, int N,
                    int rows_per_cta) {
  const uint32_t thr = rng::threshold8(p);
  const uint32_t key = thr ? rng::make_key(seed_ptr ? *seed_ptr : 0ull, op_id) : 0u;
  const float scale = rng::scale8(thr);
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(T, row0 + rows_per_cta);
  for (int c = threadIdx.x; c < N; c += 256) {
    float acc = 0.f;
    for (int r = row0; r < row1; ++r) {
      const size_t e = (size_t)r * N + c;
      float v = g[e];
      if (thr) v = rng::keep8(key, (uint32_t)e, thr) ? v * scale : 0.f;
      out[e] = __float2bfloat16(v);
      acc += v;
    }
    if (colsum) atomicAdd(colsum + c, acc);
  }
}

// 4 columns per thread (16-byte loads, 8-byte stores), N/4 lanes per row, 256/(N/4) rows in flight per CTA pass and
// 4 passes unrolled: the scalar kernel above ran at 40 % of the HBM roofline for lack of loads in flight
template <int U>
__global__ void __launch_bounds__(256)
dropout_grad_vec_kernel(const float *__restrict__ g, __nv_bfloat16 *__restrict__ out, float *__restrict__ colsum,
                        float p, const unsigned long long *__restrict__ seed_ptr, uint32_t op_id, int T, int N,
                        int rows_per_cta) {
  __shared__ float s_sum[1024];
  const uint32_t thr = rng::threshold8(p);
  const uint32_t key = thr ? rng::make_key(seed_ptr ? *seed_ptr : 0ull, op_id) : 0u;
  const float scale = rng::scale8(thr);
  const int lanes = N >> 2, ny = 256 / lanes;
  const int tx = threadIdx.x % lanes, ty = threadIdx.x / lanes;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(T, row0 + rows_per_cta);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (ty < ny) {
    for (int r = row0 + ty; r < row1; r += U * ny) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + u * ny;
        if (rr < row1) v[u] = __ldg(reinterpret_cast<const float4 *>(g + (size_t)rr * N) + tx);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + u * ny;
        if (rr >= row1) continue;
        float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        if (thr) {
          const uint32_t e = (uint32_t)((size_t)rr * N) + 4u * tx;      // N % 4 == 0 on this path: one hash per thread and row
          rng::drop_values<4>(f, key, e, thr, scale, true);
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&lo);
        pk.y = *reinterpret_cast<const uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(out + (size_t)rr * N + 4 * tx) = pk;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += f[q];
      }
    }
  }
  if (colsum) {   // fixed-order combine of the ny row-lanes, then one atomic per column and CTA
    if (ty < ny) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s_sum[ty * N + 4 * tx + q] = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += 256) {
      float t = 0.f;
      for (int y = 0; y < ny; ++y) t += s_sum[y * N + c];
      atomicAdd(colsum + c, t);
    }
  }
}

// ------------------------------------------------------ column sum / sumsq
// thread (tx, ty): tx owns a PAIR of adjacent columns (a warp row-read is 128/256 contiguous bytes), ty strides rows
template <typename T> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 ld(const float *p, size_t i) { return *reinterpret_cast<const float2 *>(p + i); }
};
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16 *p, size_t i) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p + i)); }
};

template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T *__restrict__ x, double *__restrict__ sum, double *__restrict__ sumsq, float *__restrict__ sum_f32,
              long long R, int C, int rows_per_cta, int lanes_x) {
  const int tx = threadIdx.x % lanes_x, ty = threadIdx.x / lanes_x, ny = 256 / lanes_x;
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (int cp = tx + blockIdx.y * lanes_x; cp < C / 2; cp += lanes_x * gridDim.y) {
    double a0 = 0.0, a1 = 0.0, q0 = 0.0, q1 = 0.0;
    for (long long r0 = row0 + ty; r0 < row1; r0 += (long long)ny * 32) {
      float pa0 = 0.f, pa1 = 0.f, pq0 = 0.f, pq1 = 0.f;
      const long long r1 = min(row1, r0 + (long long)ny * 32);
      for (long long r = r0; r < r1; r += ny) {
        const float2 v = Pair<T>::ld(x, (size_t)r * C + 2 * cp);
        pa0 += v.x; pa1 += v.y; pq0 += v.x * v.x; pq1 += v.y * v.y;
      }
      a0 += pa0; a1 += pa1; q0 += pq0; q1 += pq1;
    }
    if (sum) { atomicAdd(sum + 2 * cp, a0); atomicAdd(sum + 2 * cp + 1, a1); }
    if (sumsq) { atomicAdd(sumsq + 2 * cp, q0); atomicAdd(sumsq + 2 * cp + 1, q1); }
    if (sum_f32) { atomicAdd(sum_f32 + 2 * cp, (float)a0); atomicAdd(sum_f32 + 2 * cp + 1, (float)a1); }
  }
}

// ------------------------------------------------------------- BatchNorm1d
// stats[0..C) = sum, stats[C..2C) = sumsq (double).  Writes the folded affine (scale, shift) and (mean, rstd).
__global__ void bn_finalize_kernel(const double *__restrict__ stats, long long R, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float *__restrict__ running_mean,
                                   float *__restrict__ running_var, float momentum, float eps, int training,
                                   float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ mean_out,
                                   float *__restrict__ rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = stats[c] / (double)R;
    double v = stats[C + c] / (double)R - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unb = R > 1 ? v * (double)R / (double)(R - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const Tin *__restrict__ x, const float *__restrict__ scale, const float *__restrict__ shift,
                Tout *__restrict__ y, int relu, size_t total, int C) {
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int c = (int)(e % C);
    float v = ldf<Tin>(x, e) * scale[c] + shift[c];
    if (relu) v = fmaxf(v, 0.f);
    stf<Tout>(y, e, v);
  }
}

// phase 1 of BN backward: red[0..C) += sum dyb, red[C..2C) += sum dyb * xhat   (dyb = relu-masked dy)
template <typename Tdy, typename Tx>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const Tdy *__restrict__ dy, const Tx *__restrict__ x, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                     int relu, double *__restrict__ red, long long R, int C, int rows_per_cta, int lanes_x) {
  const int tx = threadIdx.x % lanes_x, ty = threadIdx.x / lanes_x, ny = 256 / lanes_x;
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (int cp = tx + blockIdx.y * lanes_x; cp < C / 2; cp += lanes_x * gridDim.y) {
    const int c = 2 * cp;
    const float sc0 = scale[c], sh0 = shift[c], mu0 = mean[c], rs0 = rstd[c];
    const float sc1 = scale[c + 1], sh1 = shift[c + 1], mu1 = mean[c + 1], rs1 = rstd[c + 1];
    double a0 = 0.0, a1 = 0.0, q0 = 0.0, q1 = 0.0;
    for (long long r0 = row0 + ty; r0 < row1; r0 += (long long)ny * 32) {
      float pa0 = 0.f, pa1 = 0.f, pq0 = 0.f, pq1 = 0.f;
      const long long r1 = min(row1, r0 + (long long)ny * 32);
      for (long long r = r0; r < r1; r += ny) {
        const float2 xv = Pair<Tx>::ld(x, (size_t)r * C + c);
        float2 d = Pair<Tdy>::ld(dy, (size_t)r * C + c);
        if (relu) {
          if (!(xv.x * sc0 + sh0 > 0.f)) d.x = 0.f;
          if (!(xv.y * sc1 + sh1 > 0.f)) d.y = 0.f;
        }
        pa0 += d.x; pa1 += d.y;
        pq0 += d.x * (xv.x - mu0) * rs0; pq1 += d.y * (xv.y - mu1) * rs1;
      }
      a0 += pa0; a1 += pa1; q0 += pq0; q1 += pq1;
    }
    atomicAdd(red + c, a0); atomicAdd(red + c + 1, a1);
    atomicAdd(red + C + c, q0); atomicAdd(red + C + c + 1, q1);
  }
}

// phase 2: dx = gamma*rstd * (dyb - sum_dyb/R - xhat * sum_dyb_xhat/R); CTA 0 also emits dgamma/dbeta (+=)
template <typename Tdy, typename Tx, typename Tdx>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const Tdy *__restrict__ dy, const Tx *__restrict__ x, const float *__restrict__ scale,
                    const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                    int relu, const double *__restrict__ red, Tdx *__restrict__ dx, float *__restrict__ dgamma,
                    float *__restrict__ dbeta, long long R, int C) {
  const size_t total = (size_t)R * C;
  const double invR = 1.0 / (double)R;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int c = (int)(e % C);
    const float sc = scale[c], sh = shift[c];
    const float xv = ldf<Tx>(x, e);
    float d = ldf<Tdy>(dy, e);
    if (relu && !(xv * sc + sh > 0.f)) d = 0.f;
    const float xh = (xv - mean[c]) * rstd[c];
    const float m1 = (float)(red[c] * invR), m2 = (float)(red[C + c] * invR);
    stf<Tdx>(dx, e, sc * (d - m1 - xh * m2));
  }
  if (blockIdx.x == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += 256) {
      dgamma[c] += (float)red[C + c];
      dbeta[c] += (float)red[c];
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// bf16 fast paths (C % 8 == 0).  Thread (tx, ty): tx owns 8 ADJACENT COLUMNS for the whole kernel (per-column
// parameters live in registers), ty strides the rows of the CTA's slab; 4 rows are in flight per thread
// (independent 16-byte loads) so one SM keeps ~128 KB of requests outstanding.
__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 t = __bfloat1622float2(h[q]); f[2 * q] = t.x; f[2 * q + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

struct ColMap {
  int cgx, ny, tx, ty, col;   // col = first of this thread's 8 columns
  bool active;
};
__device__ __forceinline__ ColMap col_map(int C) {
  ColMap m;
  const int cg = C / 8;
  m.cgx = min(cg, 256);
  m.ny = 256 / m.cgx;
  m.tx = threadIdx.x % m.cgx;
  m.ty = threadIdx.x / m.cgx;
  m.col = (m.tx + blockIdx.y * m.cgx) * 8;
  m.active = m.ty < m.ny && m.col < C;
  return m;
}

__global__ void __launch_bounds__(256)
bn_apply_bf16_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ scale, const float *__restrict__ shift,
                     __nv_bfloat16 *__restrict__ y, int relu, long long R, int C, int rows_per_cta) {
  const ColMap m = col_map(C);
  if (!m.active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { sc[q] = scale[m.col + q]; sh[q] = shift[m.col + q]; }
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long r = row0 + m.ty; r < row1; r += 4LL * m.ny) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (r + (long long)k * m.ny < row1) u[k] = ldg16(x + (size_t)(r + (long long)k * m.ny) * C + m.col);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r + (long long)k * m.ny >= row1) continue;
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int q = 0; q < 8; ++q) { f[q] = f[q] * sc[q] + sh[q]; if (relu) f[q] = fmaxf(f[q], 0.f); }
      *reinterpret_cast<uint4 *>(y + (size_t)(r + (long long)k * m.ny) * C + m.col) = pack8(f);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_bf16_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ x,
                         const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                         const float *__restrict__ rstd, int relu, const float *__restrict__ fm,
                         __nv_bfloat16 *__restrict__ dx, long long R, int C, int rows_per_cta) {
  const ColMap m = col_map(C);
  if (!m.active) return;
  float sc[8], sh[8], mu[8], rs[8], m1[8], m2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    sc[q] = scale[m.col + q]; sh[q] = shift[m.col + q]; mu[q] = mean[m.col + q]; rs[q] = rstd[m.col + q];
    m1[q] = fm[m.col + q]; m2[q] = fm[C + m.col + q];
  }
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  for (long long r = row0 + m.ty; r < row1; r += 2LL * m.ny) {
    uint4 ud[2], ux[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long rr = r + (long long)k * m.ny;
      if (rr < row1) { ud[k] = ldg16(dy + (size_t)rr * C + m.col); ux[k] = ldg16(x + (size_t)rr * C + m.col); }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long rr = r + (long long)k * m.ny;
      if (rr >= row1) continue;
      float d[8], xv[8];
      unpack8(ud[k], d);
      unpack8(ux[k], xv);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float dd = d[q];
        if (relu && !(xv[q] * sc[q] + sh[q] > 0.f)) dd = 0.f;
        d[q] = sc[q] * (dd - m1[q] - (xv[q] - mu[q]) * rs[q] * m2[q]);
      }
      *reinterpret_cast<uint4 *>(dx + (size_t)rr * C + m.col) = pack8(d);
    }
  }
}

// BN backward apply fused with the per-group row sum that follows it in Group2Emb's backward (the per-patch half of the
// split conv3 needs sum_{rows of a patch} dx, functional.group2emb_bwd): one WARP owns whole groups of S consecutive rows
// of a 256-channel tensor -- a row is exactly one 512-byte warp access, lane l owns channels 8l .. 8l+7 -- so the group sums
// stay in the lane's registers: no shared memory, no barrier, and the 1.07 GB dx tensor is not read again by a second kernel.
//   dx = a*dyb + b*x + c  with a = scale, b = -scale*rstd*m2, c = scale*(rstd*m2*mean - m1)   (same algebra as above)
__global__ void __launch_bounds__(256)
bn_bwd_apply_gsum_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ x,
                         const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                         const float *__restrict__ rstd, int relu, const float *__restrict__ fm,
                         __nv_bfloat16 *__restrict__ dx, __nv_bfloat16 *__restrict__ gsum_bf16, float *__restrict__ gsum_f32,
                         long long G, int S) {
  constexpr int C = 256;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8;
  float sc[8], sh[8], cb[8], cc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane * 8 + q;
    sc[q] = scale[c]; sh[q] = shift[c];
    const float t = sc[q] * rstd[c] * fm[C + c];
    cb[q] = -t;
    cc[q] = fmaf(t, mean[c], -sc[q] * fm[c]);
  }
  for (long long g = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); g < G; g += nwarps) {
    const size_t base = (size_t)g * S * C + lane * 8;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int r0 = 0; r0 < S; r0 += 4) {
      uint4 ud[4], ux[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r0 + k < S) { ud[k] = ldg16(dy + base + (size_t)(r0 + k) * C); ux[k] = ldg16(x + base + (size_t)(r0 + k) * C); }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (r0 + k >= S) continue;
        float d[8], xv[8];
        unpack8(ud[k], d);
        unpack8(ux[k], xv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float dd = d[q];
          if (relu && !(fmaf(xv[q], sc[q], sh[q]) > 0.f)) dd = 0.f;
          d[q] = fmaf(sc[q], dd, fmaf(cb[q], xv[q], cc[q]));
          acc[q] += d[q];
        }
        *reinterpret_cast<uint4 *>(dx + base + (size_t)(r0 + k) * C) = pack8(d);
      }
    }
    const size_t o = (size_t)g * C + lane * 8;
    if (gsum_bf16) *reinterpret_cast<uint4 *>(gsum_bf16 + o) = pack8(acc);
    if (gsum_f32) {
      *reinterpret_cast<float4 *>(gsum_f32 + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4 *>(gsum_f32 + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// column sums of a bf16 matrix: fp32 partials per thread, shared-memory combine across ty, one fp64 (or fp32) atomic
// per column per CTA.  MODE 0: sum (+ sumsq); MODE 1: BN backward phase 1 (sum dyb, sum dyb*xhat)
template <int MODE>
__global__ void __launch_bounds__(256)
colreduce_bf16_kernel(const __nv_bfloat16 *__restrict__ a, const __nv_bfloat16 *__restrict__ xin,
                      const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                      const float *__restrict__ rstd, int relu, double *__restrict__ out0, double *__restrict__ out1,
                      float *__restrict__ out0_f32, long long R, int C, int rows_per_cta) {
  // [2][ny][cgx*8]: one slot per thread, summed in a fixed order below, so that the forward BatchNorm statistics
  // (and with them the whole forward pass) are bitwise reproducible up to the fp64 atomics across CTAs
  extern __shared__ float s_red[];
  const ColMap m = col_map(C);
  const int ncol_cta = m.cgx * 8;
  if (!m.active && m.ty < m.ny) {
#pragma unroll
    for (int q = 0; q < 8; ++q) s_red[m.ty * ncol_cta + m.tx * 8 + q] = s_red[(m.ny + m.ty) * ncol_cta + m.tx * 8 + q] = 0.f;
  }
  if (m.active) {
    float sc[8], sh[8], mu[8], rs[8];
    if (MODE == 1) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { sc[q] = scale[m.col + q]; sh[q] = shift[m.col + q]; mu[q] = mean[m.col + q]; rs[q] = rstd[m.col + q]; }
    }
    float a0[8], a1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a0[q] = a1[q] = 0.f;
    const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
    constexpr int U = MODE == 1 ? 2 : 4;
    for (long long r = row0 + m.ty; r < row1; r += (long long)U * m.ny) {
      uint4 ua[U], ux[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long rr = r + (long long)k * m.ny;
        if (rr < row1) {
          ua[k] = ldg16(a + (size_t)rr * C + m.col);
          if (MODE == 1) ux[k] = ldg16(xin + (size_t)rr * C + m.col);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long rr = r + (long long)k * m.ny;
        if (rr >= row1) continue;
        float f[8];
        unpack8(ua[k], f);
        if (MODE == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) { a0[q] += f[q]; a1[q] += f[q] * f[q]; }
        } else {
          float xv[8];
          unpack8(ux[k], xv);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float d = f[q];
            if (relu && !(xv[q] * sc[q] + sh[q] > 0.f)) d = 0.f;
            a0[q] += d;
            a1[q] += d * (xv[q] - mu[q]) * rs[q];
          }
        }
      }
    }
    float4 *p0 = reinterpret_cast<float4 *>(s_red + m.ty * ncol_cta + m.tx * 8);
    float4 *p1 = reinterpret_cast<float4 *>(s_red + (m.ny + m.ty) * ncol_cta + m.tx * 8);
    p0[0] = make_float4(a0[0], a0[1], a0[2], a0[3]); p0[1] = make_float4(a0[4], a0[5], a0[6], a0[7]);
    p1[0] = make_float4(a1[0], a1[1], a1[2], a1[3]); p1[1] = make_float4(a1[4], a1[5], a1[6], a1[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncol_cta; i += 256) {
    const int c = blockIdx.y * ncol_cta + i;
    if (c >= C) continue;
    float t0 = 0.f, t1 = 0.f;
    for (int y = 0; y < m.ny; ++y) { t0 += s_red[y * ncol_cta + i]; t1 += s_red[(m.ny + y) * ncol_cta + i]; }
    if (out0) atomicAdd(out0 + c, (double)t0);
    if (out1) atomicAdd(out1 + c, (double)t1);
    if (out0_f32) atomicAdd(out0_f32 + c, t0);
  }
}

// h = gelu(z), 8 bf16 per thread (nn.GELU of the MLP, partseg.py:196).  Runs at full occupancy, which hides the
// MUFU/FMA latency chains that a 16-warp GEMM epilogue cannot.
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const uint4 *__restrict__ z, uint4 *__restrict__ h, size_t n8) {
  constexpr int U = 4;      // 4 independent 16-byte loads per thread before the first use (64 B in flight per thread)
  for (size_t i0 = (size_t)blockIdx.x * (256 * U) + threadIdx.x; i0 < n8; i0 += (size_t)gridDim.x * (256 * U)) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (i0 + (size_t)k * 256 < n8) u[k] = __ldg(z + i0 + (size_t)k * 256);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (i0 + (size_t)k * 256 >= n8) continue;
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = gelu_f(f[q]);
      h[i0 + (size_t)k * 256] = pack8(f);
    }
  }
}

// dz = dh * gelu'(z);  colsum += column sums of dz (the bias gradient of the first MLP Linear)
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16 *__restrict__ dh, const __nv_bfloat16 *__restrict__ z, __nv_bfloat16 *__restrict__ dz,
                float *__restrict__ colsum, long long R, int C, int rows_per_cta) {
  extern __shared__ float s_red[];   // [cgx*8]
  const ColMap m = col_map(C);
  const int ncol_cta = m.cgx * 8;
  for (int i = threadIdx.x; i < ncol_cta; i += 256) s_red[i] = 0.f;
  __syncthreads();
  if (m.active) {
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
    for (long long r = row0 + m.ty; r < row1; r += 2LL * m.ny) {
      uint4 ud[2], uz[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const long long rr = r + (long long)k * m.ny;
        if (rr < row1) { ud[k] = ldg16(dh + (size_t)rr * C + m.col); uz[k] = ldg16(z + (size_t)rr * C + m.col); }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const long long rr = r + (long long)k * m.ny;
        if (rr >= row1) continue;
        float d[8], zz[8];
        unpack8(ud[k], d);
        unpack8(uz[k], zz);
#pragma unroll
        for (int q = 0; q < 8; ++q) { d[q] *= gelu_grad_f(zz[q]); acc[q] += d[q]; }
        *reinterpret_cast<uint4 *>(dz + (size_t)rr * C + m.col) = pack8(d);
      }
    }
    if (colsum) {
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(&s_red[m.tx * 8 + q], acc[q]);
    }
  }
  __syncthreads();
  if (colsum) {
    for (int i = threadIdx.x; i < ncol_cta; i += 256) {
      const int c = blockIdx.y * ncol_cta + i;
      if (c < C) atomicAdd(colsum + c, s_red[i]);
    }
  }
}

// fm[0..C) = sum_dyb / R, fm[C..2C) = sum_dyb_xhat / R (fp32, for the apply pass); dgamma/dbeta accumulate
__global__ void bn_bwd_means_kernel(const double *__restrict__ red, float *__restrict__ fm, float *__restrict__ dgamma,
                                    float *__restrict__ dbeta, long long R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  fm[c] = (float)(red[c] / (double)R);
  fm[C + c] = (float)(red[C + c] / (double)R);
  if (dgamma) { dgamma[c] += (float)red[C + c]; dbeta[c] += (float)red[c]; }
}

static inline void bf16_col_cfg(long long R, int C, dim3 &grid, int &rows_per_cta, int &smem, int slots = 0) {
  const int cg = C / 8, cgx = min(cg, 256), ny = 256 / cgx, gy = ceil_div(cg, cgx);
  const long long want_ctas = max(1, (slots > 0 ? slots : num_sms() * 6) / gy);
  rows_per_cta = (int)max((long long)ny * 8, ceil_div(R, want_ctas));
  grid = dim3((unsigned)ceil_div(R, (long long)rows_per_cta), gy);
  smem = 2 * ny * cgx * 8 * (int)sizeof(float);   // colreduce: one slot per thread and statistic
}

__global__ void cast_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) y[i] = __float2bfloat16(x[i]);
}

static inline int grid_for(size_t total) { return (int)min((size_t)num_sms() * 8, ceil_div(total, (size_t)256)); }

}  // namespace vpf

using namespace vpf;
typedef __nv_bfloat16 bf16;

#define LN_DISPATCH(D, MACRO)                                                                       \
  switch (D) {                                                                                      \
    case 64: MACRO(2, false); break; case 128: MACRO(4, true); break; case 256: MACRO(8, true); break; \
    case 384: MACRO(12, true); break; case 512: MACRO(16, true); break; case 768: MACRO(24, true); break; \
    case 1024: MACRO(32, true); break;                                                              \
    default: return fail(VPF_EINVAL, "layernorm: D=%d unsupported (64,128,256,384,512,768,1024)", D); \
  }

extern "C" {

int vpf_layernorm_fwd(const void *x, int x_bf16, const float *add, int add_rows, float *xsum, const float *gamma,
                      const float *beta, void *y_bf16, float *mean, float *rstd, int T, int D, float eps, int relu,
                      void *stream) {
  VPF_REQUIRE(x && gamma && beta && y_bf16, "layernorm_fwd: null pointer");
  VPF_REQUIRE(!add || add_rows > 0, "layernorm_fwd: add_rows must be > 0");
  if (T == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(T, 8);   // 8 rows per CTA: a row-striding resident grid with the next row prefetched measured slower
                                     // (64 registers -> half the warps per SM: 256 vs 213 us on the 1 M-row kv_norm)
#define LNF_ROWS(TIN, NP, VEC)                                                                                      \
  {                                                                                                                 \
    VPF_RESIDENT_CTAS(slots, (ln_fwd_rows_kernel<TIN, NP, VEC>), 256, 0);                                            \
    ln_fwd_rows_kernel<TIN, NP, VEC><<<min(grid, slots), 256, 0, st>>>((const TIN *)x, add, add_rows, xsum, gamma, beta, (bf16 *)y_bf16, mean, rstd, T, eps, relu); \
  }
#define LNF(NP, VEC)                                                                                                \
  if (NP <= 4) { if (x_bf16) LNF_ROWS(bf16, NP, VEC) else LNF_ROWS(float, NP, VEC) }                                \
  else if (x_bf16) ln_fwd_kernel<bf16, NP, VEC><<<grid, 256, 0, st>>>((const bf16 *)x, add, add_rows, xsum, gamma, beta, (bf16 *)y_bf16, mean, rstd, T, eps, relu); \
  else ln_fwd_kernel<float, NP, VEC><<<grid, 256, 0, st>>>((const float *)x, add, add_rows, xsum, gamma, beta, (bf16 *)y_bf16, mean, rstd, T, eps, relu)
  LN_DISPATCH(D, LNF)
#undef LNF
#undef LNF_ROWS
  return check_launch("ln_fwd_kernel");
}

static int layernorm_bwd_impl(const void *dy, int dy_bf16, const void *x, int x_bf16, const void *y_relu, const float *mean,
                              const float *rstd, const float *gamma, const float *dres, void *dx, int dx_bf16, float *dgamma,
                              float *dbeta, float *dpos, int pos_rows, int T, int D, const LnEmit *emit, void *stream) {
  VPF_REQUIRE(dy && x && mean && rstd && gamma && dx, "layernorm_bwd: null pointer");
  VPF_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma/dbeta must both be given or both null");
  if (T == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tok_major = dpos && pos_rows < T;
  if (tok_major) VPF_REQUIRE(pos_rows > 0 && T % pos_rows == 0, "layernorm_bwd: T=%d is not a multiple of pos_rows=%d", T, pos_rows);
  LnEmit em{nullptr, nullptr, nullptr, 0u, 0.f};
  if (emit) {
    em = *emit;
    VPF_REQUIRE(em.g && dy_bf16 && !x_bf16 && !dx_bf16 && D % 128 == 0 && !y_relu,
                "layernorm_bwd_emit: needs bf16 dy, fp32 x / dx, D %% 128 == 0 (D=%d)", D);
    VPF_REQUIRE(em.p >= 0.f && em.p < 1.f && (size_t)T * (size_t)D < (1ull << 32), "layernorm_bwd_emit: p=%f or index space out of range", em.p);
  }
  // grid = the resident set (a partial second wave cost a third of the bandwidth).  Token-major order (see the kernel):
  // warps = pos_rows * nchunk, every (token, batch-chunk) pair one warp
#define LNB_LAUNCH(KERNEL)                                                                                            \
  {                                                                                                                   \
    VPF_RESIDENT_CTAS(slots, (KERNEL), 256, 0);                                                                       \
    int grid = min(ceil_div(T, 8), slots);                                                                            \
    const int rows_per_cta = ceil_div(T, grid);                                                                       \
    grid = ceil_div(T, rows_per_cta);                                                                                 \
    if (tok_major) {                                                                                                  \
      const int nchunk = max(1, min(T / pos_rows, (slots * 8) / pos_rows));                                           \
      grid = ceil_div(pos_rows * nchunk, 8);                                                                          \
    }                                                                                                                 \
    KERNEL<<<grid, 256, 0, st>>>((decltype(lnb_dy))dy, (decltype(lnb_x))x, (const bf16 *)y_relu, mean, rstd, gamma, dres,  \
                                 (decltype(lnb_dx))dx, dgamma, dbeta, dpos, pos_rows, T, rows_per_cta, em);          \
  }
#define LNB_CALL(TDY, TX, TDX, NP, VEC)                                                                               \
  {                                                                                                                   \
    const TDY *lnb_dy = nullptr; const TX *lnb_x = nullptr; TDX *lnb_dx = nullptr;                                    \
    (void)lnb_dy; (void)lnb_x; (void)lnb_dx;                                                                          \
    LNB_LAUNCH((ln_bwd_kernel<TDY, TX, TDX, NP, VEC, false>))                                                         \
  }
#define LNB_CALL_EMIT(NP)                                                                                             \
  {                                                                                                                   \
    const bf16 *lnb_dy = nullptr; const float *lnb_x = nullptr; float *lnb_dx = nullptr;                              \
    (void)lnb_dy; (void)lnb_x; (void)lnb_dx;                                                                          \
    LNB_LAUNCH((ln_bwd_kernel<bf16, float, float, NP, true, true>))                                                   \
  }
#define LNB(NP, VEC)                                                                     \
  if (emit) { if (VEC) LNB_CALL_EMIT(NP >= 4 ? NP : 4) }                                 \
  else if (dy_bf16 && x_bf16 && dx_bf16) LNB_CALL(bf16, bf16, bf16, NP, VEC)             \
  else if (dy_bf16 && !x_bf16 && dx_bf16) LNB_CALL(bf16, float, bf16, NP, VEC)           \
  else if (!dy_bf16 && !x_bf16 && !dx_bf16) LNB_CALL(float, float, float, NP, VEC)       \
  else if (dy_bf16 && !x_bf16 && !dx_bf16) LNB_CALL(bf16, float, float, NP, VEC)         \
  else if (dy_bf16 && x_bf16 && !dx_bf16) LNB_CALL(bf16, bf16, float, NP, VEC)           \
  else return fail(VPF_EINVAL, "layernorm_bwd: unsupported dtype combination dy_bf16=%d x_bf16=%d dx_bf16=%d", dy_bf16, x_bf16, dx_bf16)
  LN_DISPATCH(D, LNB)
#undef LNB
#undef LNB_CALL
#undef LNB_CALL_EMIT
#undef LNB_LAUNCH
  return check_launch("ln_bwd_kernel");
}

int vpf_layernorm_bwd(const void *dy, int dy_bf16, const void *x, int x_bf16, const void *y_relu, const float *mean,
                      const float *rstd, const float *gamma, const float *dres, void *dx, int dx_bf16, float *dgamma,
                      float *dbeta, float *dpos, int pos_rows, int T, int D, void *stream) {
  return layernorm_bwd_impl(dy, dy_bf16, x, x_bf16, y_relu, mean, rstd, gamma, dres, dx, dx_bf16, dgamma, dbeta, dpos, pos_rows,
                            T, D, nullptr, stream);
}

int vpf_layernorm_bwd_emit(const void *dy_bf16, const float *x, const float *mean, const float *rstd, const float *gamma,
                           const float *dres, float *dx, float *dgamma, float *dbeta, float *dpos, int pos_rows, int T, int D,
                           void *g_bf16, float *g_colsum, float drop_p, const unsigned long long *seed_ptr,
                           unsigned int op_id, void *stream) {
  VPF_REQUIRE(g_bf16, "layernorm_bwd_emit: null pointer");
  const LnEmit em{(__nv_bfloat16 *)g_bf16, g_colsum, seed_ptr, op_id, drop_p};
  return layernorm_bwd_impl(dy_bf16, 1, x, 0, nullptr, mean, rstd, gamma, dres, dx, 0, dgamma, dbeta, dpos, pos_rows, T, D, &em,
                            stream);
}

int vpf_dropout_grad(const float *g, void *out_bf16, float *colsum, float p, const unsigned long long *seed_ptr,
                     unsigned int op_id, int T, int N, void *stream) {
  VPF_REQUIRE(g && out_bf16, "dropout_grad: null pointer");
  VPF_REQUIRE(p >= 0.f && p < 1.f, "dropout_grad: p=%f out of range", p);
  VPF_REQUIRE((size_t)T * (size_t)N < (1ull << 32), "dropout_grad: index space exceeds 2^32");
  if (T == 0 || N == 0) return VPF_OK;
  VPF_RESIDENT_CTAS(slots, dropout_grad_vec_kernel<4>, 256, 0);
  const int rows_per_cta = max(1, ceil_div(T, slots));
  const int lanes = N / 4;
  if (N % 4 == 0 && lanes <= 256 && 256 % lanes == 0 && (256 / lanes) * N <= 1024 &&
      ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0) {
    dropout_grad_vec_kernel<4><<<ceil_div(T, rows_per_cta), 256, 0, (cudaStream_t)stream>>>(g, (bf16 *)out_bf16, colsum, p, seed_ptr, op_id, T, N, rows_per_cta);
    return check_launch("dropout_grad_vec_kernel");
  }
  dropout_grad_kernel<<<ceil_div(T, rows_per_cta), 256, 0, (cudaStream_t)stream>>>(g, (bf16 *)out_bf16, colsum, p, seed_ptr, op_id, T, N, rows_per_cta);
  return check_launch("dropout_grad_kernel");
}

static inline void col_launch_cfg(long long R, int C, int &lanes_x, dim3 &grid, int &rows_per_cta) {
  const int cpairs = C / 2;
  lanes_x = 1;
  while (lanes_x * 2 <= min(cpairs, 256)) lanes_x *= 2;            // power of two <= min(C/2, 256)
  const int gy = ceil_div(cpairs, lanes_x);
  rows_per_cta = (int)max((long long)(256 / lanes_x) * 32, ceil_div(R, (long long)max(1, num_sms() * 2 / gy)));
  grid = dim3((unsigned)ceil_div(R, (long long)rows_per_cta), gy);
}

int vpf_colsum(const void *x, int x_bf16, double *sum, double *sumsq, float *sum_f32, long long R, int C, void *stream) {
  VPF_REQUIRE(x && (sum || sumsq || sum_f32), "colsum: null pointer");
  VPF_REQUIRE(C % 2 == 0, "colsum: C=%d must be even", C);
  if (R == 0 || C == 0) return VPF_OK;
  if (x_bf16 && C % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 g8; int rpc, smem;
    bf16_col_cfg(R, C, g8, rpc, smem);
    VPF_RESIDENT_CTAS(slots, colreduce_bf16_kernel<0>, 256, (size_t)smem);
    bf16_col_cfg(R, C, g8, rpc, smem, slots);
    colreduce_bf16_kernel<0><<<g8, 256, smem, (cudaStream_t)stream>>>((const bf16 *)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, sum, sumsq, sum_f32, R, C, rpc);
    return check_launch("colreduce_bf16_kernel<0>");
  }
  int lanes_x, rows_per_cta;
  dim3 grid;
  col_launch_cfg(R, C, lanes_x, grid, rows_per_cta);
  if (x_bf16) colsum_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16 *)x, sum, sumsq, sum_f32, R, C, rows_per_cta, lanes_x);
  else colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)x, sum, sumsq, sum_f32, R, C, rows_per_cta, lanes_x);
  return check_launch("colsum_kernel");
}

int vpf_bn_finalize(const double *stats, long long R, const float *gamma, const float *beta, float *running_mean,
                    float *running_var, float momentum, float eps, int training, float *scale, float *shift,
                    float *mean, float *rstd, int C, void *stream) {
  VPF_REQUIRE(gamma && beta && scale && shift && mean && rstd, "bn_finalize: null pointer");
  VPF_REQUIRE(training ? (stats != nullptr && R > 0) : (running_mean && running_var), "bn_finalize: missing statistics");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, R, gamma, beta, running_mean, running_var, momentum, eps, training, scale, shift, mean, rstd, C);
  return check_launch("bn_finalize_kernel");
}

int vpf_bn_apply(const void *x, int x_bf16, const float *scale, const float *shift, void *y, int y_bf16, int relu,
                 long long R, int C, void *stream) {
  VPF_REQUIRE(x && scale && shift && y, "bn_apply: null pointer");
  const size_t total = (size_t)R * C;
  if (total == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(total);
  if (x_bf16 && y_bf16 && C % 8 == 0) {
    dim3 g8; int rpc, smem;
    VPF_RESIDENT_CTAS(slots, bn_apply_bf16_kernel, 256, 0);
    bf16_col_cfg(R, C, g8, rpc, smem, slots);
    bn_apply_bf16_kernel<<<g8, 256, 0, st>>>((const bf16 *)x, scale, shift, (bf16 *)y, relu, R, C, rpc);
  }
  else if (x_bf16 && y_bf16) bn_apply_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16 *)x, scale, shift, (bf16 *)y, relu, total, C);
  else if (!x_bf16 && y_bf16) bn_apply_kernel<float, bf16><<<grid, 256, 0, st>>>((const float *)x, scale, shift, (bf16 *)y, relu, total, C);
  else if (!x_bf16 && !y_bf16) bn_apply_kernel<float, float><<<grid, 256, 0, st>>>((const float *)x, scale, shift, (float *)y, relu, total, C);
  else bn_apply_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16 *)x, scale, shift, (float *)y, relu, total, C);
  return check_launch("bn_apply_kernel");
}

int vpf_bn_bwd(const void *dy, int dy_bf16, const void *x, int x_bf16, const float *scale, const float *shift,
               const float *mean, const float *rstd, int relu, double *red /*[3C] scratch, zeroed by the callee*/, void *dx,
               int dx_bf16, float *dgamma, float *dbeta, long long R, int C, void *stream) {
  VPF_REQUIRE(dy && x && scale && shift && mean && rstd && red && dx, "bn_bwd: null pointer");
  if (R == 0 || C == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VPF_CUDA_TRY(cudaMemsetAsync(red, 0, sizeof(double) * 2 * C, st));
  VPF_REQUIRE(C % 2 == 0, "bn_bwd: C=%d must be even", C);
  int lanes_x, rows_per_cta;
  dim3 grid;
  col_launch_cfg(R, C, lanes_x, grid, rows_per_cta);
  const int g2 = grid_for((size_t)R * C);
#define BNB(TDY, TX, TDX)                                                                                                   \
  {                                                                                                                         \
    bn_bwd_reduce_kernel<TDY, TX><<<grid, 256, 0, st>>>((const TDY *)dy, (const TX *)x, scale, shift, mean, rstd, relu, red, R, C, rows_per_cta, lanes_x); \
    VPF_TRY(check_launch("bn_bwd_reduce_kernel"));                                                                          \
    bn_bwd_apply_kernel<TDY, TX, TDX><<<g2, 256, 0, st>>>((const TDY *)dy, (const TX *)x, scale, shift, mean, rstd, relu, red, (TDX *)dx, dgamma, dbeta, R, C); \
  }
  if (dy_bf16 && x_bf16 && dx_bf16 && C % 8 == 0) {
    dim3 g8; int rpc, smem;
    bf16_col_cfg(R, C, g8, rpc, smem);
    VPF_RESIDENT_CTAS(slots, colreduce_bf16_kernel<1>, 256, (size_t)smem);
    bf16_col_cfg(R, C, g8, rpc, smem, slots);
    colreduce_bf16_kernel<1><<<g8, 256, smem, st>>>((const bf16 *)dy, (const bf16 *)x, scale, shift, mean, rstd, relu, red, red + C, nullptr, R, C, rpc);
    VPF_TRY(check_launch("colreduce_bf16_kernel<1>"));
    float *fm = reinterpret_cast<float *>(red + 2 * C);   // caller provides 3C doubles of scratch
    bn_bwd_means_kernel<<<ceil_div(C, 128), 128, 0, st>>>(red, fm, dgamma, dbeta, R, C);
    VPF_TRY(check_launch("bn_bwd_means_kernel"));
    VPF_RESIDENT_CTAS(slots2, bn_bwd_apply_bf16_kernel, 256, 0);
    bf16_col_cfg(R, C, g8, rpc, smem, slots2);
    bn_bwd_apply_bf16_kernel<<<g8, 256, 0, st>>>((const bf16 *)dy, (const bf16 *)x, scale, shift, mean, rstd, relu, fm, (bf16 *)dx, R, C, rpc);
  } else
  if (dy_bf16 && x_bf16 && dx_bf16) BNB(bf16, bf16, bf16)
  else if (!dy_bf16 && !x_bf16 && !dx_bf16) BNB(float, float, float)
  else if (dy_bf16 && !x_bf16 && dx_bf16) BNB(bf16, float, bf16)
  else if (!dy_bf16 && x_bf16 && dx_bf16) BNB(float, bf16, bf16)
  else if (dy_bf16 && !x_bf16 && !dx_bf16) BNB(bf16, float, float)
  else if (!dy_bf16 && !x_bf16 && dx_bf16) BNB(float, float, bf16)
  else return fail(VPF_EINVAL, "bn_bwd: unsupported dtype combination");
#undef BNB
  return check_launch("bn_bwd_apply_kernel");
}

int vpf_bn_bwd_gsum(const void *dy_bf16, const void *x_bf16, const float *scale, const float *shift, const float *mean,
                    const float *rstd, int relu, double *red, void *dx_bf16, float *dgamma, float *dbeta, long long R, int C,
                    int S, void *gsum_bf16, float *gsum_f32, void *stream) {
  VPF_REQUIRE(dy_bf16 && x_bf16 && scale && shift && mean && rstd && red && dx_bf16 && (gsum_bf16 || gsum_f32), "bn_bwd_gsum: null pointer");
  VPF_REQUIRE(C == 256 && S >= 1 && R % S == 0, "bn_bwd_gsum: needs C == 256 and R %% S == 0 (C=%d S=%d)", C, S);
  if (R == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VPF_CUDA_TRY(cudaMemsetAsync(red, 0, sizeof(double) * 2 * C, st));
  dim3 g8; int rpc, smem;
  bf16_col_cfg(R, C, g8, rpc, smem);
  VPF_RESIDENT_CTAS(slots, colreduce_bf16_kernel<1>, 256, (size_t)smem);
  bf16_col_cfg(R, C, g8, rpc, smem, slots);
  colreduce_bf16_kernel<1><<<g8, 256, smem, st>>>((const bf16 *)dy_bf16, (const bf16 *)x_bf16, scale, shift, mean, rstd, relu, red, red + C, nullptr, R, C, rpc);
  VPF_TRY(check_launch("colreduce_bf16_kernel<1>"));
  float *fm = reinterpret_cast<float *>(red + 2 * C);
  bn_bwd_means_kernel<<<ceil_div(C, 128), 128, 0, st>>>(red, fm, dgamma, dbeta, R, C);
  VPF_TRY(check_launch("bn_bwd_means_kernel"));
  VPF_RESIDENT_CTAS(slots2, bn_bwd_apply_gsum_kernel, 256, 0);
  const long long G = R / S;
  const int grid = (int)min((long long)slots2, ceil_div(G, 8LL));
  bn_bwd_apply_gsum_kernel<<<grid, 256, 0, st>>>((const bf16 *)dy_bf16, (const bf16 *)x_bf16, scale, shift, mean, rstd, relu, fm, (bf16 *)dx_bf16, (bf16 *)gsum_bf16, gsum_f32, G, S);
  return check_launch("bn_bwd_apply_gsum_kernel");
}

int vpf_gelu_fwd(const void *z_bf16, void *h_bf16, long long n, void *stream) {
  VPF_REQUIRE(z_bf16 && h_bf16, "gelu_fwd: null pointer");
  VPF_REQUIRE(n % 8 == 0, "gelu_fwd: n=%lld must be a multiple of 8", n);
  if (n == 0) return VPF_OK;
  VPF_RESIDENT_CTAS(slots, gelu_fwd_kernel, 256, 0);
  const int grid = (int)min((size_t)slots, ceil_div((size_t)n / 8, (size_t)1024));
  gelu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4 *)z_bf16, (uint4 *)h_bf16, (size_t)n / 8);
  return check_launch("gelu_fwd_kernel");
}

int vpf_gelu_bwd(const void *dh_bf16, const void *z_bf16, void *dz_bf16, float *colsum, long long R, int C, void *stream) {
  VPF_REQUIRE(dh_bf16 && z_bf16 && dz_bf16, "gelu_bwd: null pointer");
  VPF_REQUIRE(C % 8 == 0, "gelu_bwd: C=%d must be a multiple of 8", C);
  if (R == 0 || C == 0) return VPF_OK;
  dim3 g8; int rpc, smem;
  smem = min(C / 8, 256) * 8 * (int)sizeof(float);
  VPF_RESIDENT_CTAS(slots, gelu_bwd_kernel, 256, smem);
  bf16_col_cfg(R, C, g8, rpc, smem, slots);
  smem = min(C / 8, 256) * 8 * (int)sizeof(float);
  gelu_bwd_kernel<<<g8, 256, smem, (cudaStream_t)stream>>>((const bf16 *)dh_bf16, (const bf16 *)z_bf16, (bf16 *)dz_bf16, colsum, R, C, rpc);
  return check_launch("gelu_bwd_kernel");
}

int vpf_cast_bf16(const float *x, void *y_bf16, long long n, void *stream) {
  VPF_REQUIRE(x && y_bf16, "cast_bf16: null pointer");
  if (n == 0) return VPF_OK;
  cast_bf16_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, (bf16 *)y_bf16, (size_t)n);
  return check_launch("cast_bf16_kernel");
}

int vpf_fill_zero(void *p, long long bytes, void *stream) {
  VPF_REQUIRE(p || bytes == 0, "fill_zero: null pointer");
  if (bytes == 0) return VPF_OK;
  VPF_CUDA_TRY(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
  return VPF_OK;
}

}  // extern "C"
