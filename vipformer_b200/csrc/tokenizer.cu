// tokenizer.cu -- point-cloud tokenizer kernels for sm_100a.
//
//   K1 fps_kernel        farthest-point sampling, one CTA per cloud, the cloud
//                        resident in registers (running distance) and shared
//                        memory (centroid lookup); argmax by two redux.sync
//                        per level, one __syncthreads per iteration.
//   K2 knn_group_kernel  kNN selection, one warp per query over a shared-memory
//                        copy of the cloud: pass 1 finds a threshold that at
//                        least 32 points pass (the 32nd smallest of the 64
//                        lane-local two smallest distances), pass 2 compacts
//                        the ~40 candidates under it, a rank count orders them
//                        (streaming sorted insertion is kept as the fallback);
//                        fused with the patch gather and the reference's
//                        slot-0..2 centre subtraction; the [B,G,N] distance
//                        matrix the reference materialises never exists.
//
// Reference: vipformer/model/pointcloud/utils.py:6-141.  Arithmetic is pinned
// exactly as oracle/tokenizer_oracle.c states it (explicit __fmul_rn/__fadd_rn
// where the reference does not fuse, fmaf where it does), so indices are
// bit-exact with the oracle, including under ties (lowest index wins).
#include "common.cuh"

namespace vpf {

constexpr int kFpsThreads = 256;
constexpr int kKnnThreads = 256;
constexpr int kMaxPoints = 8192;
constexpr unsigned kFull = 0xffffffffu;

// monotone float -> uint32 map (handles the slightly negative distances the
// expanded form produces)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// ---------------------------------------------------------------- K1: FPS ---
// utils.py:56-85.  PPT = points per thread (N <= PPT * kFpsThreads).
template <int PPT>
__global__ void __launch_bounds__(kFpsThreads)
fps_kernel(const float *__restrict__ pts, int N, int C, int npoint,
           const int64_t *__restrict__ start_idx, int64_t *__restrict__ out_idx,
           float *__restrict__ centers) {
  extern __shared__ float s_xyz[];  // [N*3] AoS copy of the cloud's coordinates
  __shared__ uint32_t s_val[2][kFpsThreads / 32];
  __shared__ uint32_t s_idx[2][kFpsThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float *p = pts + (size_t)b * N * C;

  if (C == 3) {
    for (int i = tid; i < N * 3; i += kFpsThreads) s_xyz[i] = p[i];  // coalesced
  } else {
    for (int i = tid; i < N * 3; i += kFpsThreads) s_xyz[i] = p[(size_t)(i / 3) * C + (i % 3)];
  }
  __syncthreads();

  float x[PPT], y[PPT], z[PPT], run[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int i = tid + k * kFpsThreads;
    const bool ok = i < N;
    x[k] = ok ? s_xyz[3 * i + 0] : 0.f;
    y[k] = ok ? s_xyz[3 * i + 1] : 0.f;
    z[k] = ok ? s_xyz[3 * i + 2] : 0.f;
    run[k] = ok ? 1e10f : -1.0f;  // utils.py:69; padding can never win
  }

  long long st = start_idx[b];
  int far = (int)(st < 0 ? 0 : (st >= N ? N - 1 : st));

  for (int it = 0; it < npoint; ++it) {
    if (tid == 0 && out_idx) out_idx[(size_t)b * npoint + it] = far;  // utils.py:75
    if (centers && tid < C) centers[((size_t)b * npoint + it) * C + tid] = p[(size_t)far * C + tid];
    const float cx = s_xyz[3 * far + 0], cy = s_xyz[3 * far + 1], cz = s_xyz[3 * far + 2];
    float best = -1.0f;
    int besti = 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const float dx = __fsub_rn(x[k], cx), dy = __fsub_rn(y[k], cy), dz = __fsub_rn(z[k], cz);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // utils.py:79
      const float r = fminf(run[k], d);                                                                 // utils.py:81
      run[k] = r;
      if (r > best) { best = r; besti = tid + k * kFpsThreads; }  // ascending index => strict > keeps the lowest
    }
    // argmax across the CTA (utils.py:83): max value, then min index among the maxima
    const uint32_t key = best < 0.f ? 0u : __float_as_uint(best) + 1u;  // run >= +0 => bits monotone
    const uint32_t wmax = __reduce_max_sync(kFull, key);
    const uint32_t wi = __reduce_min_sync(kFull, key == wmax ? (uint32_t)besti : 0xffffffffu);
    const int buf = it & 1;
    if (lane == 0) { s_val[buf][warp] = wmax; s_idx[buf][warp] = wi; }
    __syncthreads();
    const uint32_t v = lane < kFpsThreads / 32 ? s_val[buf][lane] : 0u;
    const uint32_t vi = lane < kFpsThreads / 32 ? s_idx[buf][lane] : 0xffffffffu;
    const uint32_t m = __reduce_max_sync(kFull, v);
    far = (int)__reduce_min_sync(kFull, v == m ? vi : 0xffffffffu);
  }
}

// ------------------------------------------------- K2: kNN + gather (fused) ---
// utils.py:107-141 (selection) and utils.py:22-36 (gather + slot quirk).
// grid = (B, splits); warp w of split y handles queries q = y*per + w, +8, ...
constexpr int kKnnCap = 128;   // candidate slots per warp (two-pass selection); more candidates => streaming fallback

// expanded-form distance, arithmetic order pinned to oracle/tokenizer_oracle.c (utils.py:138-140)
__device__ __forceinline__ float knn_dist(const float4 P, float cx, float cy, float cz, float c2) {
  const float dot = fmaf(cz, P.z, fmaf(cy, P.y, __fmul_rn(cx, P.x)));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), c2), P.w);
}

// ascending bitonic sort of one float per lane
__device__ __forceinline__ float warp_sort_asc(float v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(kFull, v, j);
      const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0 || k == 32);
      v = keep_min ? fminf(v, o) : fmaxf(v, o);
    }
  }
  return v;
}

// Streaming selection: lane l keeps the l-th smallest key (ordered distance bits, index) seen so far; candidates that
// beat the current worst are inserted by ballot + shuffle.  Returns this lane's index (lane < nsample).
__device__ __forceinline__ uint32_t knn_select_stream(const float4 *s_pt, int N, int nsample, float cx, float cy, float cz,
                                                      float c2, int lane) {
  uint32_t Lhi = 0xff800000u /* f2ord(+inf) */, Llo = 0xffffffffu;
  float worst = __int_as_float(0x7f800000);
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool valid = i < N;
    const float d = knn_dist(s_pt[valid ? i : 0], cx, cy, cz, c2);
    unsigned mask = __ballot_sync(kFull, valid && d <= worst);
    if (mask) {
      const uint32_t khi = f2ord(d), klo = (uint32_t)i;
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t h = __shfl_sync(kFull, khi, src), l = __shfl_sync(kFull, klo, src);
        const bool less = (Lhi < h) || (Lhi == h && Llo < l);
        const int pos = __popc(__ballot_sync(kFull, less));
        if (pos < nsample) {
          const uint32_t uhi = __shfl_up_sync(kFull, Lhi, 1), ulo = __shfl_up_sync(kFull, Llo, 1);
          if (lane == pos) { Lhi = h; Llo = l; }
          else if (lane > pos) { Lhi = uhi; Llo = ulo; }
        }
      }
      worst = ord2f(__shfl_sync(kFull, Lhi, nsample - 1));
    }
  }
  return Llo;
}

// Two-pass selection.  Pass 1: every lane tracks the two smallest distances among its N/32 points; the 32nd smallest
// of those 64 values (bitonic half-cleaner of the two lane-sorted sequences) is a threshold tau that at least 32
// points pass, and on average only ~40 do.  Pass 2 recomputes the distances (bit-identical instruction sequence),
// compacts the points with d <= tau into shared memory as (ordered distance bits, index) keys, and a rank count over
// the candidates yields the nsample smallest in (distance, index) order -- the same total order as the streaming
// selection and as the oracle.  Returns false when the candidate list overflows (heavy ties): caller falls back.
__device__ __forceinline__ bool knn_select_twopass(const float4 *s_pt, int N, int nsample, float cx, float cy, float cz,
                                                   float c2, int lane, unsigned long long *cand, uint32_t *sel,
                                                   uint32_t &Llo) {
  const float inf = __int_as_float(0x7f800000);
  float m1 = inf, m2 = inf;
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    if (i < N) {
      const float d = knn_dist(s_pt[i], cx, cy, cz, c2);
      const float t = fmaxf(m1, d);
      m1 = fminf(m1, d);
      m2 = fminf(m2, t);
    }
  }
  const float a = warp_sort_asc(m1, lane), bs = warp_sort_asc(m2, lane);
  const float lo32 = fminf(a, __shfl_sync(kFull, bs, 31 - lane));       // the 32 smallest of the 64 values (as a set)
  const float tau = ord2f(__reduce_max_sync(kFull, f2ord(lo32)));        // ... and the largest of them
  int cnt = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool valid = i < N;
    const float d = knn_dist(s_pt[valid ? i : 0], cx, cy, cz, c2);
    const bool pred = valid && d <= tau;
    const unsigned mask = __ballot_sync(kFull, pred);
    if (mask) {
      if (pred) {
        const int pos = cnt + __popc(mask & lt);
        if (pos < kKnnCap) cand[pos] = ((unsigned long long)f2ord(d) << 32) | (uint32_t)i;
      }
      cnt += __popc(mask);
    }
  }
  if (cnt > kKnnCap) return false;     // warp-uniform
  __syncwarp();
  // rank of each candidate among all candidates; lane owns candidates lane, lane + 32, ...
  for (int j0 = 0; j0 < cnt; j0 += 64) {
    const int ja = j0 + lane, jb = j0 + 32 + lane;
    const unsigned long long ca = ja < cnt ? cand[ja] : ~0ull, cb = jb < cnt ? cand[jb] : ~0ull;
    int ra = 0, rb = 0;
    for (int m = 0; m < cnt; ++m) {
      const unsigned long long x = cand[m];     // broadcast read
      ra += x < ca;
      rb += x < cb;
    }
    if (ja < cnt && ra < nsample) sel[ra] = (uint32_t)ca;
    if (jb < cnt && rb < nsample) sel[rb] = (uint32_t)cb;
  }
  __syncwarp();
  Llo = lane < nsample ? sel[lane] : 0xffffffffu;
  __syncwarp();                        // sel aliases the gather staging buffer
  return true;
}

// Register-blocked two-pass selection: one warp handles QB queries at once, so every point is read from shared memory
// ONCE per pass for QB distance evaluations.  (One query per warp was shared-memory-bandwidth bound: 2 x 16 B per
// (query, point) pair = 8.4 MB per cloud through a 128 B/clk port.)  Same arithmetic, thresholds, candidate order and
// rank rule as knn_select_twopass, per query.  ok[q] = false when that query's candidate list overflowed.
template <int QB>
__device__ __forceinline__ void knn_select_twopass_multi(const float4 *s_pt, int N, int nsample, const float (&cx)[QB],
                                                         const float (&cy)[QB], const float (&cz)[QB], const float (&c2)[QB],
                                                         int lane, unsigned long long *cand /*[QB][kKnnCap]*/,
                                                         uint32_t *sel /*[QB][32]*/, uint32_t (&Llo)[QB], bool (&ok)[QB]) {
  const float inf = __int_as_float(0x7f800000);
  float m1[QB], m2[QB], tau[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) m1[q] = m2[q] = inf;
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    if (i < N) {
      const float4 P = s_pt[i];
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        const float d = knn_dist(P, cx[q], cy[q], cz[q], c2[q]);
        const float t = fmaxf(m1[q], d);
        m1[q] = fminf(m1[q], d);
        m2[q] = fminf(m2[q], t);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    const float a = warp_sort_asc(m1[q], lane), bs = warp_sort_asc(m2[q], lane);
    const float lo32 = fminf(a, __shfl_sync(kFull, bs, 31 - lane));
    tau[q] = ord2f(__reduce_max_sync(kFull, f2ord(lo32)));
  }
  int cnt[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) cnt[q] = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool valid = i < N;
    const float4 P = s_pt[valid ? i : 0];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      const float d = knn_dist(P, cx[q], cy[q], cz[q], c2[q]);
      const bool pred = valid && d <= tau[q];
      const unsigned mask = __ballot_sync(kFull, pred);
      if (mask) {
        if (pred) {
          const int pos = cnt[q] + __popc(mask & lt);
          if (pos < kKnnCap) cand[q * kKnnCap + pos] = ((unsigned long long)f2ord(d) << 32) | (uint32_t)i;
        }
        cnt[q] += __popc(mask);
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    ok[q] = cnt[q] <= kKnnCap;     // warp-uniform
    if (!ok[q]) continue;
    const unsigned long long *cq = cand + q * kKnnCap;
    uint32_t *sq = sel + q * 32;
    for (int j0 = 0; j0 < cnt[q]; j0 += 64) {
      const int ja = j0 + lane, jb = j0 + 32 + lane;
      const unsigned long long ca = ja < cnt[q] ? cq[ja] : ~0ull, cb = jb < cnt[q] ? cq[jb] : ~0ull;
      int ra = 0, rb = 0;
      for (int m = 0; m < cnt[q]; ++m) {
        const unsigned long long x = cq[m];     // broadcast read
        ra += x < ca;
        rb += x < cb;
      }
      if (ja < cnt[q] && ra < nsample) sq[ra] = (uint32_t)ca;
      if (jb < cnt[q] && rb < nsample) sq[rb] = (uint32_t)cb;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < QB; ++q) Llo[q] = (ok[q] && lane < nsample) ? sel[q * 32 + lane] : 0xffffffffu;
  __syncwarp();
}

constexpr int kKnnQB = 4;    // queries per warp pass

__global__ void __launch_bounds__(kKnnThreads)
knn_group_kernel(const float *__restrict__ pts, int N, int C,
                 const float *__restrict__ queries, int Q, int Cq, int nsample,
                 int q_per_cta, int64_t *__restrict__ knn_idx,
                 float *__restrict__ neighbors) {
  extern __shared__ float4 s_pt[];  // [N] (x, y, z, |p|^2)
  __shared__ float s_out[kKnnThreads / 32][32 * 3];
  __shared__ uint32_t s_sel[kKnnThreads / 32][kKnnQB * 32];
  __shared__ unsigned long long s_cand[kKnnThreads / 32][kKnnQB * kKnnCap];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float *p = pts + (size_t)b * N * C;
  for (int i = tid; i < N; i += kKnnThreads) {
    const float x = p[(size_t)i * C + 0], y = p[(size_t)i * C + 1], z = p[(size_t)i * C + 2];
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));  // utils.py:140
    s_pt[i] = make_float4(x, y, z, n2);
  }
  __syncthreads();

  const int q_begin = blockIdx.y * q_per_cta;
  const int q_end = min(Q, q_begin + q_per_cta);

  for (int q0 = q_begin + warp * kKnnQB; q0 < q_end; q0 += (kKnnThreads / 32) * kKnnQB) {
    float cx[kKnnQB], cy[kKnnQB], cz[kKnnQB], c2[kKnnQB];
#pragma unroll
    for (int j = 0; j < kKnnQB; ++j) {
      const int q = min(q0 + j, q_end - 1);      // a ragged tail repeats the last query (its outputs are skipped)
      const float *c = queries + ((size_t)b * Q + q) * Cq;
      cx[j] = c[0]; cy[j] = c[1]; cz[j] = c[2];
      c2[j] = __fadd_rn(__fadd_rn(__fmul_rn(cx[j], cx[j]), __fmul_rn(cy[j], cy[j])), __fmul_rn(cz[j], cz[j]));  // utils.py:139
    }
    uint32_t Lq[kKnnQB];   // lane l: index of the l-th nearest point by (distance, index), per query
    bool ok[kKnnQB];
    if (N >= 64) {
      knn_select_twopass_multi<kKnnQB>(s_pt, N, nsample, cx, cy, cz, c2, lane, s_cand[warp], s_sel[warp], Lq, ok);
    } else {
#pragma unroll
      for (int j = 0; j < kKnnQB; ++j) ok[j] = false;
    }
#pragma unroll
    for (int j = 0; j < kKnnQB; ++j) {
      const int q = q0 + j;
      if (q >= q_end) continue;                  // warp-uniform
      uint32_t Llo = Lq[j];
      if (!ok[j]) Llo = knn_select_stream(s_pt, N, nsample, cx[j], cy[j], cz[j], c2[j], lane);   // heavy ties / tiny clouds
      const size_t row = (size_t)b * Q + q;
      if (knn_idx && lane < nsample) knn_idx[row * nsample + lane] = (int64_t)Llo;
      if (neighbors) {
        if (C == 3) {
          if (lane < nsample) {
            const float4 P = s_pt[Llo];
            const bool sub = lane < 3;  // utils.py:36 slices the SLOT axis: only slots 0,1,2 are centred
            s_out[warp][lane * 3 + 0] = sub ? __fsub_rn(P.x, cx[j]) : P.x;
            s_out[warp][lane * 3 + 1] = sub ? __fsub_rn(P.y, cy[j]) : P.y;
            s_out[warp][lane * 3 + 2] = sub ? __fsub_rn(P.z, cz[j]) : P.z;
          }
          __syncwarp();
          float *o = neighbors + row * nsample * 3;
          for (int t = lane; t < nsample * 3; t += 32) o[t] = s_out[warp][t];  // coalesced
          __syncwarp();
        } else {
          const float *c = queries + row * Cq;
          float *o = neighbors + row * nsample * C;
          for (int t0 = 0; t0 < nsample * C; t0 += 32) {
            const int t = t0 + lane;
            const int s2 = min(t / C, nsample - 1), ch = t % C;
            const uint32_t jj = __shfl_sync(kFull, Llo, s2);
            if (t < nsample * C) {
              float v = p[(size_t)jj * C + ch];
              if (s2 < 3) v = __fsub_rn(v, c[ch]);
              o[t] = v;
            }
          }
        }
      }
    }
  }
}

// utils.py:122-141, materialised (API parity only; the fused path never calls it)
__global__ void square_distance_kernel(const float *__restrict__ src, int S, int Cs,
                                       const float *__restrict__ dst, int N, int Cd,
                                       float *__restrict__ out) {
  const int b = blockIdx.z, s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float *c = src + ((size_t)b * S + s) * Cs;
  const float *p = dst + ((size_t)b * N + i) * Cd;
  const float cx = c[0], cy = c[1], cz = c[2], x = p[0], y = p[1], z = p[2];
  const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
  const float p2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float dot = fmaf(cz, z, fmaf(cy, y, __fmul_rn(cx, x)));
  out[((size_t)b * S + s) * N + i] = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), c2), p2);
}

// utils.py:88-104
__global__ void index_points_kernel(const float *__restrict__ points, int N, int C,
                                    const int64_t *__restrict__ idx, int S, float *__restrict__ out,
                                    size_t total) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ch = (int)(t % C);
  const size_t bs = t / C;
  const size_t b = bs / S;
  long long j = idx[bs];
  j = j < 0 ? 0 : (j >= N ? N - 1 : j);
  out[t] = points[(b * N + (size_t)j) * C + ch];
}

// ------------------------------------------------------------ host launchers
static int launch_fps(const float *pts, int B, int N, int C, int npoint, const int64_t *start,
                      int64_t *out_idx, float *centers, cudaStream_t st) {
  if (B == 0 || npoint == 0) return VPF_OK;
  const size_t smem = (size_t)N * 3 * sizeof(float);
#define VPF_FPS_CASE(PPT)                                                                         \
  if (N <= PPT * kFpsThreads) {                                                                   \
    if (smem > 48 * 1024)                                                                         \
      VPF_CUDA_TRY(cudaFuncSetAttribute(fps_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    fps_kernel<PPT><<<B, kFpsThreads, smem, st>>>(pts, N, C, npoint, start, out_idx, centers);    \
    return check_launch("fps_kernel");                                                            \
  }
  VPF_FPS_CASE(1) VPF_FPS_CASE(2) VPF_FPS_CASE(4) VPF_FPS_CASE(8) VPF_FPS_CASE(10) VPF_FPS_CASE(16) VPF_FPS_CASE(32)
#undef VPF_FPS_CASE
  return fail(VPF_EINVAL, "fps: N=%d exceeds the supported maximum %d", N, kMaxPoints);
}

static int launch_knn(const float *pts, int B, int N, int C, const float *queries, int Q, int Cq,
                      int nsample, int64_t *knn_idx, float *neighbors, cudaStream_t st) {
  if (B == 0 || Q == 0) return VPF_OK;
  const size_t smem = (size_t)N * sizeof(float4);
  if (smem + 40 * 1024 > 48 * 1024)   // the kernel also holds ~37 KB of static shared memory (candidates, selections, gather staging)
    VPF_CUDA_TRY(cudaFuncSetAttribute(knn_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // enough CTAs for >= 2 waves on small batches; a split re-stages the cloud (cheap)
  const int warps = (kKnnThreads / 32) * kKnnQB;     // queries one CTA handles per pass over the cloud
  int splits = ceil_div(2 * num_sms() * 4, B);
  splits = max(1, min(splits, ceil_div(Q, warps)));
  const int per = ceil_div(ceil_div(Q, splits), warps) * warps;
  splits = ceil_div(Q, per);
  knn_group_kernel<<<dim3(B, splits), kKnnThreads, smem, st>>>(pts, N, C, queries, Q, Cq, nsample, per,
                                                               knn_idx, neighbors);
  return check_launch("knn_group_kernel");
}

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_fps(const float *pts, int B, int N, int C, int npoint, const int64_t *start_idx,
            int64_t *out_idx, void *stream) {
  VPF_REQUIRE(pts && start_idx && out_idx, "fps: null pointer");
  VPF_REQUIRE(B >= 0 && N >= 1 && N <= kMaxPoints && C >= 3 && npoint >= 0, "fps: bad shape B=%d N=%d C=%d npoint=%d", B, N, C, npoint);
  return launch_fps(pts, B, N, C, npoint, start_idx, out_idx, nullptr, (cudaStream_t)stream);
}

int vpf_index_points(const float *points, int B, int N, int C, const int64_t *idx, int S, float *out,
                     void *stream) {
  VPF_REQUIRE(points && idx && out, "index_points: null pointer");
  VPF_REQUIRE(B >= 0 && N >= 1 && C >= 1 && S >= 0, "index_points: bad shape");
  const size_t total = (size_t)B * S * C;
  if (total == 0) return VPF_OK;
  index_points_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, (cudaStream_t)stream>>>(points, N, C, idx, S, out, total);
  return check_launch("index_points_kernel");
}

int vpf_square_distance(const float *src, int B, int S, int Cs, const float *dst, int N, int Cd,
                        float *out, void *stream) {
  VPF_REQUIRE(src && dst && out, "square_distance: null pointer");
  VPF_REQUIRE(B >= 0 && S >= 0 && N >= 0 && Cs >= 3 && Cd >= 3 && S <= 65535 && B <= 65535, "square_distance: bad shape");
  if (B == 0 || S == 0 || N == 0) return VPF_OK;
  square_distance_kernel<<<dim3(ceil_div(N, 256), S, B), 256, 0, (cudaStream_t)stream>>>(src, S, Cs, dst, N, Cd, out);
  return check_launch("square_distance_kernel");
}

int vpf_knn_point(int nsample, const float *xyz, int B, int N, int C, const float *new_xyz, int S,
                  int Cq, int64_t *out_idx, void *stream) {
  VPF_REQUIRE(xyz && new_xyz && out_idx, "knn_point: null pointer");
  VPF_REQUIRE(nsample >= 1 && nsample <= 32, "knn_point: nsample=%d unsupported (1..32)", nsample);
  VPF_REQUIRE(B >= 0 && N >= nsample && N <= kMaxPoints && C >= 3 && Cq >= 3 && S >= 0, "knn_point: bad shape B=%d N=%d C=%d S=%d (need nsample <= N <= %d)", B, N, C, S, kMaxPoints);
  return launch_knn(xyz, B, N, C, new_xyz, S, Cq, nsample, out_idx, nullptr, (cudaStream_t)stream);
}

int vpf_divide_patches(const float *pts, int B, int N, int C, int G, int S, const int64_t *start_idx,
                       float *neighbors, float *centers, int64_t *fps_idx, int64_t *knn_idx, void *stream) {
  VPF_REQUIRE(pts && start_idx && neighbors && centers, "divide_patches: null pointer");
  VPF_REQUIRE(S >= 1 && S <= 32, "divide_patches: group_size=%d unsupported (1..32)", S);
  VPF_REQUIRE(B >= 0 && N >= S && N <= kMaxPoints && C >= 3 && G >= 0, "divide_patches: bad shape B=%d N=%d C=%d G=%d S=%d (need S <= N <= %d)", B, N, C, G, S, kMaxPoints);
  cudaStream_t st = (cudaStream_t)stream;
  VPF_TRY(launch_fps(pts, B, N, C, G, start_idx, fps_idx, centers, st));
  return launch_knn(pts, B, N, C, centers, G, C, S, knn_idx, neighbors, st);
}

size_t vpf_divide_patches_host_workspace_bytes(int B, int N, int C, int G, int S) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return al((size_t)B * N * C * 4) + al((size_t)B * 8) + al((size_t)B * G * S * C * 4) + al((size_t)B * G * C * 4);
}

int vpf_divide_patches_host(const float *pts_host, int B, int N, int C, int G, int S,
                            const int64_t *start_idx_host, float *neighbors_host, float *centers_host,
                            void *workspace, size_t workspace_bytes, void *stream) {
  VPF_REQUIRE(pts_host && start_idx_host && neighbors_host && centers_host && workspace, "divide_patches_host: null pointer");
  if (workspace_bytes < vpf_divide_patches_host_workspace_bytes(B, N, C, G, S))
    return fail(VPF_EWORKSPACE, "divide_patches_host: workspace %zu < %zu bytes", workspace_bytes,
                vpf_divide_patches_host_workspace_bytes(B, N, C, G, S));
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)workspace;
  float *d_pts = (float *)w; w += al((size_t)B * N * C * 4);
  int64_t *d_start = (int64_t *)w; w += al((size_t)B * 8);
  float *d_nb = (float *)w; w += al((size_t)B * G * S * C * 4);
  float *d_ce = (float *)w;
  VPF_CUDA_TRY(cudaMemcpyAsync(d_pts, pts_host, (size_t)B * N * C * 4, cudaMemcpyHostToDevice, st));
  VPF_CUDA_TRY(cudaMemcpyAsync(d_start, start_idx_host, (size_t)B * 8, cudaMemcpyHostToDevice, st));
  VPF_TRY(vpf_divide_patches(d_pts, B, N, C, G, S, d_start, d_nb, d_ce, nullptr, nullptr, stream));
  VPF_CUDA_TRY(cudaMemcpyAsync(neighbors_host, d_nb, (size_t)B * G * S * C * 4, cudaMemcpyDeviceToHost, st));
  VPF_CUDA_TRY(cudaMemcpyAsync(centers_host, d_ce, (size_t)B * G * C * 4, cudaMemcpyDeviceToHost, st));
  VPF_CUDA_TRY(cudaStreamSynchronize(st));
  return VPF_OK;
}

}  // extern "C"
