// loss_optim.cu -- NT-Xent contrastive objective (fused similarity -> online log-sum-exp ->
// reduction; the [2b, 2b] logits never exist in memory) and the fused AdamW update.
//
// NT-Xent restates lightly==1.1.21 lightly/loss/ntx_ent_loss.py (third-party, not vendored by the
// reference; call sites pretrain.py:155,196,202):  rows = cat(normalize(out0), normalize(out1)),
// logits = rows rows^T / T with the diagonal removed, label = the other view, CE mean.
//   loss = mean_i [ logsumexp_{j != i} s_ij - s_{i,pos(i)} ],  s_ij = z_i . z_j / T
// Global-negative extension: rows are this rank's 2b embeddings, columns are the
// all-gathered 2bW embeddings; local row i sits at column self(i), its positive at pos(i).
// AdamW follows torch.optim.AdamW defaults (pretrain.py:121-124).
#include "common.cuh"

namespace vpf {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// z = x / max(||x||, 1e-12)  (F.normalize, dim=1);  one warp per row
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float *__restrict__ x, float *__restrict__ z, float *__restrict__ norm, int n, int D) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = x[(size_t)row * D + d]; s += v * v; }
  const float nr = fmaxf(sqrtf(wsum(s)), 1e-12f);
  for (int d = lane; d < D; d += 32) z[(size_t)row * D + d] = x[(size_t)row * D + d] / nr;
  if (lane == 0) norm[row] = nr;
}

__device__ __forceinline__ void self_pos(int i, int b_local, int col_offset, int half, int &self, int &pos) {
  if (i < b_local) { self = col_offset + i; pos = half + col_offset + i; }
  else { self = half + col_offset + (i - b_local); pos = col_offset + (i - b_local); }
}

constexpr int kMaxDPerLane = 24;  // D <= 768

// one warp per local row, lanes across the feature dim; online log-sum-exp over all columns
__global__ void __launch_bounds__(256)
ntxent_fwd_kernel(const float *__restrict__ zr, int n_r, const float *__restrict__ zc, int n_c, int D, int b_local,
                  int col_offset, int half, float invT, float *__restrict__ lse_out, float *__restrict__ loss_out) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n_r) return;
  const int nper = D / 32;
  float a[kMaxDPerLane];
#pragma unroll
  for (int t = 0; t < kMaxDPerLane; ++t) a[t] = t < nper ? zr[(size_t)i * D + lane + 32 * t] : 0.f;
  int self, pos;
  self_pos(i, b_local, col_offset, half, self, pos);
  float m = -INFINITY, l = 0.f, spos = 0.f;
  for (int j = 0; j < n_c; ++j) {
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxDPerLane; ++t) if (t < nper) dot += a[t] * zc[(size_t)j * D + lane + 32 * t];
    const float s = wsum(dot) * invT;
    if (j == pos) spos = s;
    if (j != self) {
      const float mn = fmaxf(m, s);
      l = l * __expf(m - mn) + __expf(s - mn);
      m = mn;
    }
  }
  if (lane == 0) {
    const float lse = m + __logf(l);
    lse_out[i] = lse;
    atomicAdd(loss_out, (lse - spos) / (float)n_r);
  }
}

// gradient w.r.t. the UN-normalised local rows x_k (z_k = x_k / norm_k):
//   g_k = gscale/T * [ sum_{j != self(k)} (p_kj + p_jk) z_j - 2 z_pos(k) ],  p_kj = exp(s_kj - lse_k), p_jk = exp(s_kj - lse_j)
//   dx_k = (g_k - z_k (z_k . g_k)) / norm_k
// lse_all holds the log-sum-exp of EVERY column-as-row (all-gathered across ranks).
__global__ void __launch_bounds__(256)
ntxent_bwd_kernel(const float *__restrict__ zr, const float *__restrict__ norm, int n_r, const float *__restrict__ zc,
                  const float *__restrict__ lse_all, int n_c, int D, int b_local, int col_offset, int half, float invT,
                  float gscale, const float *__restrict__ upstream, float *__restrict__ dx) {
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= n_r) return;
  const int nper = D / 32;
  float a[kMaxDPerLane], g[kMaxDPerLane];
#pragma unroll
  for (int t = 0; t < kMaxDPerLane; ++t) { a[t] = t < nper ? zr[(size_t)k * D + lane + 32 * t] : 0.f; g[t] = 0.f; }
  int self, pos;
  self_pos(k, b_local, col_offset, half, self, pos);
  const float lse_k = lse_all[self];
  for (int j = 0; j < n_c; ++j) {
    float c[kMaxDPerLane];
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxDPerLane; ++t) {
      c[t] = t < nper ? zc[(size_t)j * D + lane + 32 * t] : 0.f;
      dot += a[t] * c[t];
    }
    const float s = wsum(dot) * invT;
    float w = 0.f;
    if (j != self) w = __expf(s - lse_k) + __expf(s - lse_all[j]);
    if (j == pos) w -= 2.f;
#pragma unroll
    for (int t = 0; t < kMaxDPerLane; ++t) g[t] += w * c[t];
  }
  const float up = upstream ? upstream[0] : 1.f;
  const float sc = gscale * invT * up;
  float zg = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxDPerLane; ++t) { g[t] *= sc; zg += a[t] * g[t]; }
  zg = wsum(zg);
  const float inv = 1.f / norm[k];
#pragma unroll
  for (int t = 0; t < kMaxDPerLane; ++t)
    if (t < nper) dx[(size_t)k * D + lane + 32 * t] = (g[t] - a[t] * zg) * inv;
}

// fused AdamW (torch.optim.AdamW semantics) over a flat parameter buffer; refreshes the bf16 shadow.
// state[0] = step count (int64, already advanced for this step), lr read from device memory.
__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
             __nv_bfloat16 *__restrict__ shadow, size_t n, const float *__restrict__ lr_ptr, float beta1, float beta2,
             float eps, float wd, const long long *__restrict__ step_ptr, float grad_scale) {
  const float lr = *lr_ptr;
  const float t = (float)(*step_ptr);
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow) shadow[i] = __float2bfloat16(pi);
  }
}

// per-step device state: state[0] += 1 (optimizer step), state[1] = next dropout seed
__global__ void step_advance_kernel(long long *state) {
  state[0] += 1;
  unsigned long long s = (unsigned long long)state[1];
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  state[1] = (long long)s;
}

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_l2norm_rows(const float *x, float *z, float *norm, int n, int D, void *stream) {
  VPF_REQUIRE(x && z && norm, "l2norm_rows: null pointer");
  if (n == 0) return VPF_OK;
  l2norm_rows_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(x, z, norm, n, D);
  return check_launch("l2norm_rows_kernel");
}

int vpf_ntxent_fwd(const float *zr, int n_r, const float *zc, int n_c, int D, int b_local, int col_offset, int half,
                   float temperature, float *lse_out, float *loss_out, void *stream) {
  VPF_REQUIRE(zr && zc && lse_out && loss_out, "ntxent_fwd: null pointer");
  VPF_REQUIRE(D % 32 == 0 && D <= 32 * kMaxDPerLane, "ntxent_fwd: D=%d unsupported (multiple of 32, <= %d)", D, 32 * kMaxDPerLane);
  VPF_REQUIRE(n_r == 2 * b_local && n_c == 2 * half && col_offset >= 0 && col_offset + b_local <= half && temperature > 0.f, "ntxent_fwd: inconsistent sizes");
  if (n_r == 0) return VPF_OK;
  ntxent_fwd_kernel<<<ceil_div(n_r, 8), 256, 0, (cudaStream_t)stream>>>(zr, n_r, zc, n_c, D, b_local, col_offset, half, 1.f / temperature, lse_out, loss_out);
  return check_launch("ntxent_fwd_kernel");
}

int vpf_ntxent_bwd(const float *zr, const float *norm, int n_r, const float *zc, const float *lse_all, int n_c, int D,
                   int b_local, int col_offset, int half, float temperature, float gscale, const float *upstream,
                   float *dx, void *stream) {
  VPF_REQUIRE(zr && norm && zc && lse_all && dx, "ntxent_bwd: null pointer");
  VPF_REQUIRE(D % 32 == 0 && D <= 32 * kMaxDPerLane, "ntxent_bwd: D=%d unsupported", D);
  VPF_REQUIRE(n_r == 2 * b_local && n_c == 2 * half && col_offset >= 0 && col_offset + b_local <= half && temperature > 0.f, "ntxent_bwd: inconsistent sizes");
  if (n_r == 0) return VPF_OK;
  ntxent_bwd_kernel<<<ceil_div(n_r, 8), 256, 0, (cudaStream_t)stream>>>(zr, norm, n_r, zc, lse_all, n_c, D, b_local, col_offset, half, 1.f / temperature, gscale, upstream, dx);
  return check_launch("ntxent_bwd_kernel");
}

int vpf_adamw(float *p, const float *g, float *m, float *v, void *shadow_bf16, long long n, const float *lr_ptr,
              float beta1, float beta2, float eps, float weight_decay, const long long *step_ptr, float grad_scale,
              void *stream) {
  VPF_REQUIRE(p && g && m && v && lr_ptr && step_ptr, "adamw: null pointer");
  if (n == 0) return VPF_OK;
  const int grid = (int)min((size_t)num_sms() * 8, ceil_div((size_t)n, (size_t)256));
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16 *)shadow_bf16, (size_t)n, lr_ptr, beta1, beta2, eps, weight_decay, step_ptr, grad_scale);
  return check_launch("adamw_kernel");
}

int vpf_step_advance(long long *state, void *stream) {
  VPF_REQUIRE(state, "step_advance: null pointer");
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  return check_launch("step_advance_kernel");
}

}  // extern "C"
