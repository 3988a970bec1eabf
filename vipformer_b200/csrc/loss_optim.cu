// loss_optim.cu -- NT-Xent contrastive objective (fp32 similarity GEMM -> masked row log-sum-exp -> reduction; the
// backward reuses the logits scratch in place) and the fused AdamW update.
//
// NT-Xent restates lightly==1.1.21 lightly/loss/ntx_ent_loss.py (third-party, not vendored by the
// reference; call sites pretrain.py:155,196,202):  rows = cat(normalize(out0), normalize(out1)),
// logits = rows rows^T / T with the diagonal removed, label = the other view, CE mean.
//   loss = mean_i [ logsumexp_{j != i} s_ij - s_{i,pos(i)} ],  s_ij = z_i . z_j / T
// Global-negative extension: rows are this rank's 2b embeddings, columns are the
// all-gathered 2bW embeddings; local row i sits at column self(i), its positive at pos(i).
// Column j of the logits lives at physical row  (j / blk) * ld + base + j % blk  of the gathered buffer (ColMap): one
// NCCL all-gather of every rank's PACKED [imid rows | cmid rows] block then serves both loss terms without a re-layout.
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

struct ColMap { int blk, ld, base; };
__device__ __forceinline__ size_t phys(const ColMap &c, int j) { return (size_t)(j / c.blk) * c.ld + c.base + (j % c.blk); }

__device__ __forceinline__ void self_pos(int i, int b_local, int col_offset, int half, int &self, int &pos) {
  if (i < b_local) { self = col_offset + i; pos = half + col_offset + i; }
  else { self = half + col_offset + (i - b_local); pos = col_offset + (i - b_local); }
}

// ---- tiled fp32 SIMT GEMM for the similarity matrix and its backward (sizes are tiny: 2b x 2bW x D).  fp32 on purpose:
// logits are divided by T = 0.1, bf16 operands would put ~1e-2 of noise on them.
//   b_nt = 1: C[M,N] = alpha * A[M,K] . B[N,K]^T        b_nt = 0: C[M,N] = alpha * A[M,K] . B[K,N]
constexpr int kSgTile = 64, kSgK = 64;   // K slab of 64: 4x fewer load -> sync -> FMA -> sync rounds than 16 (the kernel is latency bound: 32-64 CTAs)
__global__ void __launch_bounds__(256)
sgemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int M, int N, int K,
             float alpha, int b_nt, ColMap cm, int cm_base_stride) {   // cm maps the logits-column index (n if b_nt, else k) to a row of B
  __shared__ float sA[kSgK][kSgTile + 4], sB[kSgK][kSgTile + 4];
  // blockIdx.z = loss term of a packed call: its own A rows, C block and column-map base; B (all gathered rows) is shared
  A += (size_t)blockIdx.z * M * K;
  C += (size_t)blockIdx.z * M * N;
  cm.base += (int)blockIdx.z * cm_base_stride;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kSgTile, n0 = blockIdx.x * kSgTile;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kSgK) {
    for (int t = threadIdx.x; t < kSgTile * kSgK; t += 256) {
      const int r = t / kSgK, kk = t % kSgK;   // A tile: rows m, contiguous in k
      const int m = m0 + r, k = k0 + kk;
      sA[kk][r] = (m < M && k < K) ? A[(size_t)m * K + k] : 0.f;
      if (b_nt) {
        const int n = n0 + r;
        sB[kk][r] = (n < N && k < K) ? B[phys(cm, n) * K + k] : 0.f;
      }
    }
    if (!b_nt) {
      for (int t = threadIdx.x; t < kSgTile * kSgK; t += 256) {
        const int kk = t / kSgTile, c = t % kSgTile;   // B tile: rows k, contiguous in n
        const int k = k0 + kk, n = n0 + c;
        sB[kk][c] = (n < N && k < K) ? B[phys(cm, k) * N + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSgK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; bb[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[(size_t)m * N + n] = alpha * acc[i][j];
    }
  }
}

// per local row i of the logits S [n_r, n_c]: lse over j != self(i), loss += (lse - S[i,pos(i)]) / n_r ; one warp per row
__global__ void __launch_bounds__(256)
ntxent_rowlse_kernel(const float *__restrict__ S, int n_r, int n_c, int b_local, int col_offset, int half,
                     float *__restrict__ lse_out, float *__restrict__ loss_out) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n_r) return;
  S += (size_t)blockIdx.y * n_r * n_c;      // blockIdx.y = loss term of a packed call
  lse_out += (size_t)blockIdx.y * n_r;
  loss_out += blockIdx.y;
  int self, pos;
  self_pos(i, b_local, col_offset, half, self, pos);
  const float *row = S + (size_t)i * n_c;
  float m = -INFINITY;
  for (int j = lane; j < n_c; j += 32) if (j != self) m = fmaxf(m, row[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
  for (int j = lane; j < n_c; j += 32) if (j != self) l += __expf(row[j] - m);
  l = wsum(l);
  if (lane == 0) {
    const float lse = m + __logf(l);
    lse_out[i] = lse;
    atomicAdd(loss_out, (lse - row[pos]) / (float)n_r);
  }
}

// in place: S[k,j] <- w_kj = [j != self(k)] (exp(s - lse_k) + exp(s - lse_all[j])) - 2 [j == pos(k)]
__global__ void __launch_bounds__(256)
ntxent_weights_kernel(float *__restrict__ S, const float *__restrict__ lse_all, int n_r, int n_c, int b_local,
                      int col_offset, int half, ColMap cm, int cm_base_stride) {
  const size_t total = (size_t)n_r * n_c;
  S += (size_t)blockIdx.y * total;          // blockIdx.y = loss term of a packed call
  cm.base += (int)blockIdx.y * cm_base_stride;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int k = (int)(e / n_c), j = (int)(e % n_c);
    int self, pos;
    self_pos(k, b_local, col_offset, half, self, pos);
    const float s = S[e];
    float w = 0.f;
    if (j != self) w = __expf(s - lse_all[phys(cm, self)]) + __expf(s - lse_all[phys(cm, j)]);
    if (j == pos) w -= 2.f;
    S[e] = w;
  }
}

// dx_k = (g_k - z_k (z_k . g_k)) / norm_k with g_k = sc * G[k,:]   (backward of F.normalize); one warp per row
struct SegScale { float v[4]; };
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float *__restrict__ G, const float *__restrict__ z, const float *__restrict__ norm,
                  SegScale sc, int seg_rows, const float *__restrict__ upstream, float *__restrict__ dx, int n, int D) {
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= n) return;
  const float s = sc.v[k / seg_rows] * (upstream ? upstream[0] : 1.f);   // rows [t * seg_rows, +seg_rows) belong to loss term t
  float zg = 0.f;
  for (int d = lane; d < D; d += 32) zg += z[(size_t)k * D + d] * G[(size_t)k * D + d] * s;
  zg = wsum(zg);
  const float inv = 1.f / norm[k];
  for (int d = lane; d < D; d += 32) dx[(size_t)k * D + d] = (G[(size_t)k * D + d] * s - z[(size_t)k * D + d] * zg) * inv;
}

// nn.CrossEntropyLoss(label_smoothing = eps), mean reduction (ft_cls.py:145,176): one warp per sample.
//   loss_i = (1 - eps) * (-log p_i[y_i]) + eps / C * sum_c (-log p_i[c]);   dlogits_i = (p_i - ((1 - eps) onehot(y_i) + eps / C)) / n
// forward and the logit gradient in one pass (the gradient is scaled by the upstream scalar in the autograd backward).
__global__ void __launch_bounds__(256)
ce_ls_kernel(const float *__restrict__ logits, int ld, const long long *__restrict__ labels, int n, int C, float eps,
             float *__restrict__ loss_out, float *__restrict__ dlogits, int ldd) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float *row = logits + (size_t)i * ld;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float se = 0.f, sx = 0.f;
  for (int c = lane; c < C; c += 32) { se += __expf(row[c] - m); sx += row[c]; }
  se = wsum(se);
  sx = wsum(sx);
  const float lse = m + __logf(se);
  const int y = (int)labels[i];
  const float inv_n = 1.f / (float)n, u = eps / (float)C;
  for (int c = lane; c < C; c += 32) {
    const float p = __expf(row[c] - lse);
    dlogits[(size_t)i * ldd + c] = (p - ((c == y ? 1.f - eps : 0.f) + u)) * inv_n;
  }
  if (lane == 0) {
    const float nll = lse - row[y], smooth = lse - sx / (float)C;
    atomicAdd(loss_out, ((1.f - eps) * nll + eps * smooth) * inv_n);
  }
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

static int check_cols(const char *what, int n_r, int n_c, int b_local, int col_offset, int half, float temperature,
                      int zc_blk, int zc_ld) {
  VPF_REQUIRE(n_r == 2 * b_local && half >= b_local && col_offset >= 0 && col_offset + half + b_local <= n_c && temperature > 0.f,
              "%s: inconsistent sizes (n_r=%d n_c=%d b=%d col_offset=%d half=%d)", what, n_r, n_c, b_local, col_offset, half);
  VPF_REQUIRE(zc_blk >= 1 && zc_ld >= zc_blk, "%s: bad column map (blk=%d ld=%d)", what, zc_blk, zc_ld);
  return VPF_OK;
}

// Packed form: nseg (<= 4) loss terms in ONE launch per stage.  zr [nseg * n_r, D] (term t = rows [t n_r, +n_r)), S_ws
// [nseg, n_r, n_c], lse_out [nseg * n_r], loss_out [nseg]; the column map base of term t is zc_base + t * zc_base_stride.
int vpf_ntxent_pack_fwd(const float *zr, int nseg, int n_r, const float *zc, int n_c, int D, int b_local, int col_offset,
                        int half, int zc_blk, int zc_ld, int zc_base, int zc_base_stride, float temperature, float *S_ws,
                        float *lse_out, float *loss_out, void *stream) {
  VPF_REQUIRE(zr && zc && S_ws && lse_out && loss_out, "ntxent_fwd: null pointer");
  VPF_REQUIRE(nseg >= 1 && nseg <= 4, "ntxent_fwd: nseg=%d out of range (1..4)", nseg);
  VPF_TRY(check_cols("ntxent_fwd", n_r, n_c, b_local, col_offset, half, temperature, zc_blk, zc_ld));
  if (n_r == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const ColMap cm{zc_blk, zc_ld, zc_base};
  sgemm_kernel<<<dim3(ceil_div(n_c, kSgTile), ceil_div(n_r, kSgTile), nseg), 256, 0, st>>>(zr, zc, S_ws, n_r, n_c, D, 1.f / temperature, 1, cm, zc_base_stride);
  VPF_TRY(check_launch("sgemm_kernel"));
  ntxent_rowlse_kernel<<<dim3(ceil_div(n_r, 8), nseg), 256, 0, st>>>(S_ws, n_r, n_c, b_local, col_offset, half, lse_out, loss_out);
  return check_launch("ntxent_rowlse_kernel");
}

int vpf_ntxent_fwd(const float *zr, int n_r, const float *zc, int n_c, int D, int b_local, int col_offset, int half,
                   int zc_blk, int zc_ld, int zc_base, float temperature, float *S_ws, float *lse_out, float *loss_out,
                   void *stream) {
  return vpf_ntxent_pack_fwd(zr, 1, n_r, zc, n_c, D, b_local, col_offset, half, zc_blk, zc_ld, zc_base, 0, temperature, S_ws,
                             lse_out, loss_out, stream);
}

// gscale [nseg] (host array): per-term seed of the gradient; dx [nseg * n_r, D]; G_ws [nseg * n_r, D]
int vpf_ntxent_pack_bwd(const float *zr, const float *norm, int nseg, int n_r, const float *zc, const float *lse_all, int n_c,
                        int D, int b_local, int col_offset, int half, int zc_blk, int zc_ld, int zc_base, int zc_base_stride,
                        float temperature, const float *gscale, const float *upstream, float *S_ws, float *G_ws, float *dx,
                        void *stream) {
  VPF_REQUIRE(zr && norm && zc && lse_all && S_ws && G_ws && dx && gscale, "ntxent_bwd: null pointer");
  VPF_REQUIRE(nseg >= 1 && nseg <= 4, "ntxent_bwd: nseg=%d out of range (1..4)", nseg);
  VPF_TRY(check_cols("ntxent_bwd", n_r, n_c, b_local, col_offset, half, temperature, zc_blk, zc_ld));
  if (n_r == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)n_r * n_c;
  const ColMap cm{zc_blk, zc_ld, zc_base};
  const int wgrid = (int)min((size_t)max(1, num_sms() * 8 / nseg), ceil_div(total, (size_t)256));
  ntxent_weights_kernel<<<dim3(wgrid, nseg), 256, 0, st>>>(S_ws, lse_all, n_r, n_c, b_local, col_offset, half, cm, zc_base_stride);
  VPF_TRY(check_launch("ntxent_weights_kernel"));
  sgemm_kernel<<<dim3(ceil_div(D, kSgTile), ceil_div(n_r, kSgTile), nseg), 256, 0, st>>>(S_ws, zc, G_ws, n_r, D, n_c, 1.f, 0, cm, zc_base_stride);
  VPF_TRY(check_launch("sgemm_kernel"));
  SegScale sc;
  for (int t = 0; t < 4; ++t) sc.v[t] = t < nseg ? gscale[t] / temperature : 0.f;
  l2norm_bwd_kernel<<<ceil_div(nseg * n_r, 8), 256, 0, st>>>(G_ws, zr, norm, sc, n_r, upstream, dx, nseg * n_r, D);
  return check_launch("l2norm_bwd_kernel");
}

int vpf_ntxent_bwd(const float *zr, const float *norm, int n_r, const float *zc, const float *lse_all, int n_c, int D,
                   int b_local, int col_offset, int half, int zc_blk, int zc_ld, int zc_base, float temperature,
                   float gscale, const float *upstream, float *S_ws, float *G_ws, float *dx, void *stream) {
  return vpf_ntxent_pack_bwd(zr, norm, 1, n_r, zc, lse_all, n_c, D, b_local, col_offset, half, zc_blk, zc_ld, zc_base, 0,
                             temperature, &gscale, upstream, S_ws, G_ws, dx, stream);
}

int vpf_ce_ls(const float *logits, int ld, const long long *labels, int n, int C, float eps, float *loss_out,
              float *dlogits, int ldd, void *stream) {
  VPF_REQUIRE(logits && labels && loss_out && dlogits && C >= 1 && ld >= C && ldd >= C && eps >= 0.f && eps < 1.f, "ce_ls: bad arguments");
  if (n == 0) return VPF_OK;
  ce_ls_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, n, C, eps, loss_out, dlogits, ldd);
  return check_launch("ce_ls_kernel");
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
