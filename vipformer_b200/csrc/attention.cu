// attention.cu -- fused multi-head attention core (softmax(Q K^T * scale) V) forward + backward for
// the small-sequence blocks of partseg.py:67-86: Lq <= 144 query tokens, Lk in {96..144} (self /
// image) or 1024..2500 (point cross-attention), head dim 64.  Logits never touch HBM.
//
// v1 data path: flash-style tiles through shared memory with mma.sync.m16n8k16 (bf16 in, fp32
// accumulate), online softmax in the exp2 domain, attention-probability dropout (p = atten_drop,
// partseg.py:81) regenerated from (seed, op_id, element) in the backward pass.  The attention
// core is ~12 % of the model FLOPs; the 85 % that is Linear/Conv runs on tcgen05 (gemm.cu).
//
// Layout: Q rows are token rows of a [B*Lq, ldq] bf16 array, head h at columns [h*64, h*64+64);
// K/V likewise in [B*Lk, ldkv]; O in [B*Lq, ldo].  LSE is fp32 [B*H, Lq] in log2 units.
#include <stdlib.h>
#include "common.cuh"
#include "rng.cuh"

namespace vpf {

typedef __nv_bfloat16 bf16;
constexpr int HD = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

// A [rows][64] bf16 tile, 128 B per row, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// cooperative load of `rows` x 64 bf16 from global (row stride ld) into a swizzled tile; rows >= valid are zeroed
__device__ __forceinline__ void load_tile(bf16 *tile, const bf16 *g, int ld, int valid, int rows, int tid, int nthr) {
  for (int i = tid; i < rows * 8; i += nthr) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid) v = *reinterpret_cast<const uint4 *>(g + (size_t)r * ld + c * 8);
    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(tile) + tile_off(r, c)) = v;
  }
}

// same, asynchronously (cp.async, 16 B per request; rows >= valid are zero-filled through src-size 0)
__device__ __forceinline__ void load_tile_async(bf16 *tile, const bf16 *g, int ld, int valid, int rows, int tid, int nthr) {
  for (int i = tid; i < rows * 8; i += nthr) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < valid;
    const bf16 *src = ok ? g + (size_t)r * ld + c * 8 : g;
    const uint32_t dst = smem_addr(reinterpret_cast<char *>(tile) + tile_off(r, c));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A fragments (16 rows x 64 cols) of the tile rows [row0, row0+16): 4 k-steps x 4 regs
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const bf16 *tile, int row0, int lane) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm4(a[kk], smem_addr(tile) + tile_off(row0 + (m & 1) * 8 + r, 2 * kk + (m >> 1)));
}

// C[16 x 64] += A(16 x 64, frags) * T^T where T is a [64 rows][64] tile: C[i][n] = sum_d A[i][d] T[n][d]
__device__ __forceinline__ void mma_a_tileT(float (&c)[8][4], const uint32_t (&a)[4][4], const bf16 *tile, int lane) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int nb = 0; nb < 8; nb += 2) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b[4];
      ldsm4(b, smem_addr(tile) + tile_off((nb + (m >> 1)) * 8 + r, 2 * kk + (m & 1)));
      mma16816(c[nb], a[kk], b[0], b[1]);
      mma16816(c[nb + 1], a[kk], b[2], b[3]);
    }
  }
}

// C[16 x 64] += P(16 x 64, packed from accumulators) * T where T is a [64 rows][64] tile: C[i][n] = sum_j P[i][j] T[j][n]
__device__ __forceinline__ void mma_p_tile(float (&c)[8][4], const uint32_t (&p)[4][4], const bf16 *tile, int lane) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int nd = 0; nd < 8; nd += 2) {
      uint32_t b[4];
      ldsm4t(b, smem_addr(tile) + tile_off(16 * kk + (m & 1) * 8 + r, nd + (m >> 1)));
      mma16816(c[nd], p[kk], b[0], b[1]);
      mma16816(c[nd + 1], p[kk], b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void pack_p(uint32_t (&p)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    p[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    p[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    p[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    p[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

// Attention-probability dropout (partseg.py:81).  One 32-bit hash covers 4 ADJACENT KEYS of one query row, 8 bits each
// (layout shared with attention_tc.cu, where a thread owns a row segment; restated in oracle/rng.py attention_keep).
// Effective p = round(256 p) / 256 (0.1 -> 26/256 = 0.1016).
struct DropCfg {
  uint32_t thr, key;   // thr: 8-bit threshold (0 = dropout off)
  float scale;
};
__device__ __forceinline__ DropCfg make_drop(float p, const unsigned long long *seed_ptr, uint32_t op_id) {
  DropCfg d;
  d.thr = p > 0.f ? (uint32_t)(p * 256.f + 0.5f) : 0u;
  d.key = p > 0.f ? rng::make_key(seed_ptr ? *seed_ptr : 0ull, op_id) : 0u;
  d.scale = d.thr ? 256.f / (256.f - (float)d.thr) : 1.f;
  return d;
}
// hash of the 4-key group containing (i, j); element (i, j) uses byte j & 3
__device__ __forceinline__ uint32_t block_hash(const DropCfg &dc, int bh, int i, int j, int Lq, int Lk) {
  const uint32_t idx = ((uint32_t)bh * (uint32_t)Lq + (uint32_t)i) * (uint32_t)((Lk + 3) >> 2) + (uint32_t)(j >> 2);
  return rng::mix32(idx * 0x9e3779b1u ^ dc.key);
}
__device__ __forceinline__ bool keep_from(uint32_t hash, int i, int j, uint32_t thr) {
  return ((hash >> ((j & 3) * 8)) & 0xffu) >= thr;
}

// ------------------------------------------------------------------ forward
// NW warps x 16 query rows per CTA; K/V streamed in 64-key tiles, double-buffered with cp.async
template <int NW>
__global__ void __launch_bounds__(32 * NW)
attn_fwd_kernel(const bf16 *__restrict__ Q, int ldq, const bf16 *__restrict__ K, const bf16 *__restrict__ V, int ldkv,
                bf16 *__restrict__ O, int ldo, float *__restrict__ LSE, int H, int Lq, int Lk, float scale,
                float drop_p, const unsigned long long *__restrict__ seed_ptr, uint32_t op_id) {
  constexpr int TM = 16 * NW, NT = 32 * NW;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  bf16 *sQ = reinterpret_cast<bf16 *>(attn_smem);
  bf16 *sK = sQ + TM * HD;          // [2][64*HD]
  bf16 *sV = sK + 2 * 64 * HD;      // [2][64*HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * TM;
  const DropCfg dc = make_drop(drop_p, seed_ptr, op_id);
  const bf16 *Kb = K + (size_t)b * Lk * ldkv + h * HD, *Vb = V + (size_t)b * Lk * ldkv + h * HD;

  load_tile_async(sQ, Q + ((size_t)b * Lq + q0) * ldq + h * HD, ldq, Lq - q0, TM, tid, NT);
  load_tile_async(sK, Kb, ldkv, Lk, 64, tid, NT);
  load_tile_async(sV, Vb, ldkv, Lk, 64, tid, NT);
  cp_async_commit();

  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sc2 = scale * kLog2e;
  const int r0 = q0 + warp * 16 + (lane >> 2);  // this thread's first query row (second is +8)
  const int nt = (Lk + 63) / 64;

  for (int t = 0; t < nt; ++t) {
    const int k0 = t * 64;
    if (t + 1 < nt) {
      load_tile_async(sK + ((t + 1) & 1) * 64 * HD, Kb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT);
      load_tile_async(sV + ((t + 1) & 1) * 64 * HD, Vb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) load_a_frags(qa, sQ, warp * 16, lane);
    const bf16 *tK = sK + (t & 1) * 64 * HD, *tV = sV + (t & 1) * 64 * HD;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_a_tileT(s, qa, tK, lane);
    // scale, mask the key tail, running max
    float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = k0 + nb * 8 + (lane & 3) * 2 + (e & 1);
        const float v = j < Lk ? s[nb][e] * sc2 : -INFINITY;
        s[nb][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 1));
      mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      corr[q] = exp2f(mrow[q] - mx[q]);  // mrow = -inf on the first tile -> 0
      mrow[q] = mx[q];
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nb][e] - mx[e >> 1]);
        rs[e >> 1] += p;
        s[nb][e] = p;
      }
      if (dc.thr) {
        const int j = k0 + nb * 8 + (lane & 3) * 2;   // even: (j, j+1) share a hash
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int i = r0 + q * 8;
          const uint32_t hsh = block_hash(dc, bh, i, j, Lq, Lk);
          s[nb][2 * q] = keep_from(hsh, i, j, dc.thr) ? s[nb][2 * q] * dc.scale : 0.f;
          s[nb][2 * q + 1] = keep_from(hsh, i, j + 1, dc.thr) ? s[nb][2 * q + 1] * dc.scale : 0.f;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) lrow[q] = lrow[q] * corr[q] + rs[q];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      o[nd][0] *= corr[0]; o[nd][1] *= corr[0];
      o[nd][2] *= corr[1]; o[nd][3] *= corr[1];
    }
    uint32_t pa[4][4];
    pack_p(pa, s);
    mma_p_tile(o, pa, tV, lane);
    __syncthreads();   // tile buffer (t & 1) may be refilled by the prefetch of iteration t + 1
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    lrow[q] += __shfl_xor_sync(0xffffffffu, lrow[q], 1);
    lrow[q] += __shfl_xor_sync(0xffffffffu, lrow[q], 2);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = r0 + q * 8;
    if (i < Lq) {
      const float inv = 1.f / lrow[q];
      bf16 *op = O + ((size_t)b * Lq + i) * ldo + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        *reinterpret_cast<uint32_t *>(op + nd * 8) = pack_bf16(o[nd][2 * q] * inv, o[nd][2 * q + 1] * inv);
      if ((lane & 3) == 0) LSE[(size_t)bh * Lq + i] = mrow[q] + log2f(lrow[q]);
    }
  }
}

// delta[bh, i] = sum_d dO[i, d] * O[i, d]
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16 *__restrict__ O, int ldo, const bf16 *__restrict__ dO, int lddo, float *__restrict__ delta,
                  int H, int Lq, long long total) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;  // over (b, i, h)
  if (t >= total) return;
  const int h = (int)(t % H);
  const long long row = t / H;  // b*Lq + i
  const int b = (int)(row / Lq), i = (int)(row % Lq);
  const uint4 *po = reinterpret_cast<const uint4 *>(O + (size_t)row * ldo + h * HD);
  const uint4 *pd = reinterpret_cast<const uint4 *>(dO + (size_t)row * lddo + h * HD);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 a = po[c], d = pd[c];
    const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&a), *hd = reinterpret_cast<const __nv_bfloat162 *>(&d);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 fa = __bfloat1622float2(ha[q]), fd = __bfloat1622float2(hd[q]);
      acc += fa.x * fd.x + fa.y * fd.y;
    }
  }
  delta[((size_t)b * H + h) * Lq + i] = acc;
}

// ------------------------------------------------- backward: dK, dV (per key block)
// Each warp owns 16 keys (NW*16 keys per CTA) and walks all queries in chunks of 64 (double-buffered cp.async),
// working on S^T = K Q^T.
template <int NW>
__global__ void __launch_bounds__(32 * NW)
attn_bwd_dkv_kernel(const bf16 *__restrict__ Q, int ldq, const bf16 *__restrict__ K, const bf16 *__restrict__ V,
                    int ldkv, const bf16 *__restrict__ dO, int lddo, const float *__restrict__ LSE,
                    const float *__restrict__ delta, bf16 *__restrict__ dK, bf16 *__restrict__ dV, int lddkv, int H,
                    int Lq, int Lk, float scale, float drop_p, const unsigned long long *__restrict__ seed_ptr,
                    uint32_t op_id) {
  constexpr int TK = 16 * NW, NT = 32 * NW;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  bf16 *sK = reinterpret_cast<bf16 *>(attn_smem);
  bf16 *sV = sK + TK * HD;
  bf16 *sQ = sV + TK * HD;           // [2][64*HD]
  bf16 *sdO = sQ + 2 * 64 * HD;      // [2][64*HD]
  float *sLse = reinterpret_cast<float *>(sdO + 2 * 64 * HD);   // [2][64]
  float *sDelta = sLse + 2 * 64;                                 // [2][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * TK;
  const DropCfg dc = make_drop(drop_p, seed_ptr, op_id);
  const float sc2 = scale * kLog2e;
  const bf16 *Qb = Q + (size_t)b * Lq * ldq + h * HD, *dOb = dO + (size_t)b * Lq * lddo + h * HD;

  load_tile_async(sK, K + ((size_t)b * Lk + k0) * ldkv + h * HD, ldkv, Lk - k0, TK, tid, NT);
  load_tile_async(sV, V + ((size_t)b * Lk + k0) * ldkv + h * HD, ldkv, Lk - k0, TK, tid, NT);
  load_tile_async(sQ, Qb, ldq, Lq, 64, tid, NT);
  load_tile_async(sdO, dOb, lddo, Lq, 64, tid, NT);
  cp_async_commit();

  uint32_t ka[4][4], va[4][4];
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  const int jrow = k0 + warp * 16 + (lane >> 2);  // this thread's first key (second is +8)
  const int nq = (Lq + 63) / 64;

  for (int t = 0; t < nq; ++t) {
    const int q0 = t * 64, buf = t & 1;
    if (t + 1 < nq) {
      load_tile_async(sQ + (buf ^ 1) * 64 * HD, Qb + (size_t)(q0 + 64) * ldq, ldq, Lq - q0 - 64, 64, tid, NT);
      load_tile_async(sdO + (buf ^ 1) * 64 * HD, dOb + (size_t)(q0 + 64) * lddo, lddo, Lq - q0 - 64, 64, tid, NT);
    }
    cp_async_commit();
    if (tid < 64) {
      const int i = q0 + tid;
      sLse[buf * 64 + tid] = i < Lq ? LSE[(size_t)bh * Lq + i] : INFINITY;   // +inf -> p = 0 for padded queries
      sDelta[buf * 64 + tid] = i < Lq ? delta[(size_t)bh * Lq + i] : 0.f;
    }
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) {
      load_a_frags(ka, sK, warp * 16, lane);
      load_a_frags(va, sV, warp * 16, lane);
    }
    const bf16 *tQ = sQ + buf * 64 * HD, *tdO = sdO + buf * 64 * HD;
    const float *tl = sLse + buf * 64, *td = sDelta + buf * 64;
    float st[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    mma_a_tileT(st, ka, tQ, lane);    // S^T[key][q]
    mma_a_tileT(dp, va, tdO, lane);   // dP^T[key][q] = V dO^T
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int qe = nb * 8 + (lane & 3) * 2;   // even local query
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int j = jrow + h2 * 8;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int e = h2 * 2 + w, ql = qe + w;
          const float p = j < Lk ? exp2f(st[nb][e] * sc2 - tl[ql]) : 0.f;
          float dpe = dp[nb][e];
          float pd = p;
          if (dc.thr) {
            const bool keep = keep_from(block_hash(dc, bh, q0 + ql, j, Lq, Lk), q0 + ql, j, dc.thr);
            pd = keep ? p * dc.scale : 0.f;
            dpe = keep ? dpe * dc.scale : 0.f;
          }
          st[nb][e] = pd;                               // dropped P^T (for dV)
          dp[nb][e] = p * (dpe - td[ql]) * scale;       // dS^T (for dK)
        }
      }
    }
    uint32_t pa[4][4];
    pack_p(pa, st);
    mma_p_tile(dv, pa, tdO, lane);   // dV += P^T dO
    pack_p(pa, dp);
    mma_p_tile(dk, pa, tQ, lane);    // dK += dS^T Q
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = jrow + q * 8;
    if (j < Lk) {
      bf16 *pk = dK + ((size_t)b * Lk + j) * lddkv + h * HD + (lane & 3) * 2;
      bf16 *pv = dV + ((size_t)b * Lk + j) * lddkv + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        *reinterpret_cast<uint32_t *>(pk + nd * 8) = pack_bf16(dk[nd][2 * q], dk[nd][2 * q + 1]);
        *reinterpret_cast<uint32_t *>(pv + nd * 8) = pack_bf16(dv[nd][2 * q], dv[nd][2 * q + 1]);
      }
    }
  }
}

// ------------------------------------------------------ backward: dQ (per query block)
template <int NW>
__global__ void __launch_bounds__(32 * NW)
attn_bwd_dq_kernel(const bf16 *__restrict__ Q, int ldq, const bf16 *__restrict__ K, const bf16 *__restrict__ V,
                   int ldkv, const bf16 *__restrict__ dO, int lddo, const float *__restrict__ LSE,
                   const float *__restrict__ delta, bf16 *__restrict__ dQ, int lddq, int H, int Lq, int Lk,
                   float scale, float drop_p, const unsigned long long *__restrict__ seed_ptr, uint32_t op_id) {
  constexpr int TM = 16 * NW, NT = 32 * NW;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  bf16 *sQ = reinterpret_cast<bf16 *>(attn_smem);
  bf16 *sdO = sQ + TM * HD;
  bf16 *sK = sdO + TM * HD;         // [2][64*HD]
  bf16 *sV = sK + 2 * 64 * HD;      // [2][64*HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * TM;
  const DropCfg dc = make_drop(drop_p, seed_ptr, op_id);
  const float sc2 = scale * kLog2e;
  const bf16 *Kb = K + (size_t)b * Lk * ldkv + h * HD, *Vb = V + (size_t)b * Lk * ldkv + h * HD;

  load_tile_async(sQ, Q + ((size_t)b * Lq + q0) * ldq + h * HD, ldq, Lq - q0, TM, tid, NT);
  load_tile_async(sdO, dO + ((size_t)b * Lq + q0) * lddo + h * HD, lddo, Lq - q0, TM, tid, NT);
  load_tile_async(sK, Kb, ldkv, Lk, 64, tid, NT);
  load_tile_async(sV, Vb, ldkv, Lk, 64, tid, NT);
  cp_async_commit();

  uint32_t qa[4][4], da[4][4];
  const int r0 = q0 + warp * 16 + (lane >> 2);
  float lse[2], dl[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = r0 + q * 8;
    lse[q] = i < Lq ? LSE[(size_t)bh * Lq + i] : INFINITY;
    dl[q] = i < Lq ? delta[(size_t)bh * Lq + i] : 0.f;
  }
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const int nt = (Lk + 63) / 64;

  for (int t = 0; t < nt; ++t) {
    const int k0 = t * 64;
    if (t + 1 < nt) {
      load_tile_async(sK + ((t + 1) & 1) * 64 * HD, Kb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT);
      load_tile_async(sV + ((t + 1) & 1) * 64 * HD, Vb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) {
      load_a_frags(qa, sQ, warp * 16, lane);
      load_a_frags(da, sdO, warp * 16, lane);
    }
    const bf16 *tK = sK + (t & 1) * 64 * HD, *tV = sV + (t & 1) * 64 * HD;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    mma_a_tileT(s, qa, tK, lane);    // S = Q K^T
    mma_a_tileT(dp, da, tV, lane);   // dP = dO V^T
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int je = k0 + nb * 8 + (lane & 3) * 2;   // even key: (je, je+1) share a hash
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = r0 + q * 8;
        uint32_t hsh = 0;
        if (dc.thr) hsh = block_hash(dc, bh, i, je, Lq, Lk);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int e = q * 2 + w, j = je + w;
          const float p = j < Lk ? exp2f(s[nb][e] * sc2 - lse[q]) : 0.f;
          float dpe = dp[nb][e];
          if (dc.thr) dpe = keep_from(hsh, i, j, dc.thr) ? dpe * dc.scale : 0.f;
          s[nb][e] = p * (dpe - dl[q]) * scale;  // dS
        }
      }
    }
    uint32_t pa[4][4];
    pack_p(pa, s);
    mma_p_tile(dq, pa, tK, lane);    // dQ += dS K
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = r0 + q * 8;
    if (i < Lq) {
      bf16 *pq = dQ + ((size_t)b * Lq + i) * lddq + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) *reinterpret_cast<uint32_t *>(pq + nd * 8) = pack_bf16(dq[nd][2 * q], dq[nd][2 * q + 1]);
    }
  }
}

// ------------------------------------------- backward, fused (Lq <= 128): dQ, dK, dV in ONE pass
// One CTA (8 warps) per (sample, head) holds all queries.  Per 64-key tile:
//   phase 1 (warp = 16 query rows):  S = Q K^T, P = exp2(S - lse), dP = dO V^T, dS = P (dP - delta) scale;
//                                    dQ += dS K (registers); dropped P and dS go to shared memory as bf16 tiles
//   phase 2 (warp = 16 keys x 32 d): dV = P^T dO, dK = dS^T Q with the A operand read TRANSPOSED (ldmatrix.trans)
// so S / P / dP / dS and the dropout mask are computed once (the two-kernel path computes them twice: 5 GEMMs vs 7).
template <int NQW>
__global__ void __launch_bounds__(32 * NQW)
attn_bwd_fused_kernel(const bf16 *__restrict__ Q, int ldq, const bf16 *__restrict__ K, const bf16 *__restrict__ V,
                      int ldkv, const bf16 *__restrict__ O, int ldo, const bf16 *__restrict__ dO, int lddo,
                      const float *__restrict__ LSE, bf16 *__restrict__ dQ, int lddq, bf16 *__restrict__ dK,
                      bf16 *__restrict__ dV, int lddkv, int H, int Lq, int Lk, float scale, float drop_p,
                      const unsigned long long *__restrict__ seed_ptr, uint32_t op_id) {
  extern __shared__ __align__(128) uint8_t attn_smem[];
  constexpr int TQ = 16 * NQW, NT_ = 32 * NQW;
  bf16 *sQ = reinterpret_cast<bf16 *>(attn_smem);   // [TQ][64]
  bf16 *sdO = sQ + TQ * HD;                         // [TQ][64]
  bf16 *sK = sdO + TQ * HD;                         // [2][64][64]
  bf16 *sV = sK + 2 * 64 * HD;                      // [2][64][64]
  bf16 *sP = sV + 2 * 64 * HD;                      // [TQ q][64 keys]  dropped probabilities
  bf16 *sdS = sP + TQ * 64;                         // [TQ q][64 keys]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const DropCfg dc = make_drop(drop_p, seed_ptr, op_id);
  const float sc2 = scale * kLog2e;
  const bf16 *Kb = K + (size_t)b * Lk * ldkv + h * HD, *Vb = V + (size_t)b * Lk * ldkv + h * HD;

  load_tile_async(sQ, Q + (size_t)b * Lq * ldq + h * HD, ldq, Lq, TQ, tid, NT_);
  load_tile_async(sdO, dO + (size_t)b * Lq * lddo + h * HD, lddo, Lq, TQ, tid, NT_);
  load_tile_async(sK, Kb, ldkv, Lk, 64, tid, NT_);
  load_tile_async(sV, Vb, ldkv, Lk, 64, tid, NT_);
  cp_async_commit();

  uint32_t qa[4][4], da[4][4];
  const int r0 = warp * 16 + (lane >> 2);   // this thread's first query row (second is +8)
  float lse[2], dl[2];
  {
    // delta_i = sum_d dO[i,d] O[i,d] for this warp's 16 rows: two lanes per row, 32 head-dim columns each
    const int ri = warp * 16 + (lane >> 1), half = lane & 1;
    float acc = 0.f;
    if (ri < Lq) {
      const uint4 *po = reinterpret_cast<const uint4 *>(O + ((size_t)b * Lq + ri) * ldo + h * HD + half * 32);
      const uint4 *pd = reinterpret_cast<const uint4 *>(dO + ((size_t)b * Lq + ri) * lddo + h * HD + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = __ldg(po + c), d = __ldg(pd + c);
        const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&a), *hd = reinterpret_cast<const __nv_bfloat162 *>(&d);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = __bfloat1622float2(ha[q]), fd = __bfloat1622float2(hd[q]);
          acc += fa.x * fd.x + fa.y * fd.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);   // lanes 2r, 2r+1 now both hold delta of row warp*16 + r
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = r0 + q * 8;
      lse[q] = i < Lq ? LSE[(size_t)bh * Lq + i] : INFINITY;
      dl[q] = __shfl_sync(0xffffffffu, acc, 2 * ((lane >> 2) + q * 8));
    }
  }
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const int kb = (warp & 3) * 16, dh = (warp >> 2) * 32;   // phase-2 ownership: 16 keys x 32 head-dim columns
  const int m = lane >> 3, rr = lane & 7;
  const int nt = (Lk + 63) / 64;

  for (int t = 0; t < nt; ++t) {
    const int k0 = t * 64;
    if (t + 1 < nt) {
      load_tile_async(sK + ((t + 1) & 1) * 64 * HD, Kb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT_);
      load_tile_async(sV + ((t + 1) & 1) * 64 * HD, Vb + (size_t)(k0 + 64) * ldkv, ldkv, Lk - k0 - 64, 64, tid, NT_);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();   // tile t landed; phase 2 of tile t-1 is finished with sP / sdS
    if (t == 0) {
      load_a_frags(qa, sQ, warp * 16, lane);
      load_a_frags(da, sdO, warp * 16, lane);
    }
    const bf16 *tK = sK + (t & 1) * 64 * HD, *tV = sV + (t & 1) * 64 * HD;
    {
      // ---------------- phase 1
      float s[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      mma_a_tileT(s, qa, tK, lane);    // S = Q K^T
      mma_a_tileT(dp, da, tV, lane);   // dP = dO V^T
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int je = k0 + nb * 8 + (lane & 3) * 2;   // even key: (je, je+1) share a dropout hash
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int i = r0 + q * 8;
          uint32_t hsh = 0;
          if (dc.thr) hsh = block_hash(dc, bh, i, je, Lq, Lk);
          float pd[2], ds[2];
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int e = q * 2 + w, j = je + w;
            const float p = j < Lk ? exp2f(s[nb][e] * sc2 - lse[q]) : 0.f;
            float dpe = dp[nb][e];
            pd[w] = p;
            if (dc.thr) {
              const bool keep = keep_from(hsh, i, j, dc.thr);
              pd[w] = keep ? p * dc.scale : 0.f;
              dpe = keep ? dpe * dc.scale : 0.f;
            }
            ds[w] = p * (dpe - dl[q]) * scale;
            s[nb][e] = ds[w];
          }
          const uint32_t off = tile_off(warp * 16 + (lane >> 2) + q * 8, nb) + (lane & 3) * 4;
          *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(sP) + off) = pack_bf16(pd[0], pd[1]);
          *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(sdS) + off) = pack_bf16(ds[0], ds[1]);
        }
      }
      uint32_t pa[4][4];
      pack_p(pa, s);
      mma_p_tile(dq, pa, tK, lane);    // dQ += dS K
    }
    __syncthreads();   // sP / sdS complete
    if (warp < 8) {
      // ---------------- phase 2: dV[kb..+16][dh..+32] = P^T dO, dK = dS^T Q  (k dimension = the TQ queries)
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NQW; ++kk) {
        uint32_t ap[4], as[4];
        const uint32_t aoff = tile_off(16 * kk + (m >> 1) * 8 + rr, kb / 8 + (m & 1));
        ldsm4t(ap, smem_addr(sP) + aoff);
        ldsm4t(as, smem_addr(sdS) + aoff);
#pragma unroll
        for (int nd = 0; nd < 4; nd += 2) {
          uint32_t bo[4], bq[4];
          const uint32_t boff = tile_off(16 * kk + (m & 1) * 8 + rr, dh / 8 + nd + (m >> 1));
          ldsm4t(bo, smem_addr(sdO) + boff);
          ldsm4t(bq, smem_addr(sQ) + boff);
          mma16816(dv[nd], ap, bo[0], bo[1]);
          mma16816(dv[nd + 1], ap, bo[2], bo[3]);
          mma16816(dk[nd], as, bq[0], bq[1]);
          mma16816(dk[nd + 1], as, bq[2], bq[3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = k0 + kb + (lane >> 2) + q * 8;
        if (j < Lk) {
          bf16 *pk = dK + ((size_t)b * Lk + j) * lddkv + h * HD + dh + (lane & 3) * 2;
          bf16 *pv = dV + ((size_t)b * Lk + j) * lddkv + h * HD + dh + (lane & 3) * 2;
#pragma unroll
          for (int nd = 0; nd < 4; ++nd) {
            *reinterpret_cast<uint32_t *>(pk + nd * 8) = pack_bf16(dk[nd][2 * q], dk[nd][2 * q + 1]);
            *reinterpret_cast<uint32_t *>(pv + nd * 8) = pack_bf16(dv[nd][2 * q], dv[nd][2 * q + 1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = r0 + q * 8;
    if (i < Lq) {
      bf16 *pq = dQ + ((size_t)b * Lq + i) * lddq + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) *reinterpret_cast<uint32_t *>(pq + nd * 8) = pack_bf16(dq[nd][2 * q], dq[nd][2 * q + 1]);
    }
  }
}

// rows per CTA: 128 (8 warps) when that pads no more than 64-row tiles would, else 64 (4 warps)
static inline bool use8(int L) { return ceil_div(L, 128) * 128 <= ceil_div(L, 64) * 64; }
template <typename Kern> static int set_smem(Kern k, int bytes) {
  if (bytes > 48 * 1024) VPF_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return VPF_OK;
}

// tcgen05 / TMEM path (attention_tc.cu): every attention with Lq <= 256 query tokens (point-cloud AND image branch)
namespace atc {
int attention_tc_fwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, void *O, int ldo, float *LSE, int B,
                     int H, int Lq, int Lk, float scale, float drop_p, const unsigned long long *seed_ptr,
                     unsigned int op_id, cudaStream_t st);
int attention_tc_bwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, const void *O, int ldo, const void *dO,
                     int lddo, const float *LSE, void *dQ, int lddq, void *dK, void *dV, int lddkv, int B, int H, int Lq,
                     int Lk, float scale, float drop_p, const unsigned long long *seed_ptr, unsigned int op_id,
                     cudaStream_t st);
}  // namespace atc
static bool use_tc(int Lq, float drop_p) {
  static const int off = [] { const char *v = getenv("VPF_ATTN_TC"); return (v && v[0] == '0') ? 1 : 0; }();   // experiments: VPF_ATTN_TC=0
  // two query blocks at most (two dQ accumulators in TMEM); dropout threshold <= 128 / 256 (byte-parallel compare)
  return !off && Lq <= 256 && drop_p <= 0.5f;
}
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_attention_fwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, void *O, int ldo, float *LSE,
                      int B, int H, int Lq, int Lk, int head_dim, float scale, float drop_p,
                      const unsigned long long *seed_ptr, unsigned int op_id, void *stream) {
  VPF_REQUIRE(Q && K && V && O && LSE, "attention_fwd: null pointer");
  VPF_REQUIRE(head_dim == HD, "attention_fwd: head_dim=%d unsupported (64)", head_dim);
  VPF_REQUIRE(Lq >= 1 && Lk >= 1 && H >= 1 && (ldq % 8) == 0 && (ldkv % 8) == 0 && (ldo % 8) == 0, "attention_fwd: bad shape/stride");
  VPF_REQUIRE((long long)B * H <= 65535 * 1LL * 65535, "attention_fwd: too many heads");
  if (B == 0) return VPF_OK;
  if (use_tc(Lq, drop_p) && aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O))
    return atc::attention_tc_fwd(Q, ldq, K, V, ldkv, O, ldo, LSE, B, H, Lq, Lk, scale, drop_p, seed_ptr, op_id, (cudaStream_t)stream);
  if (Lq > 128 && Lq <= 160) {   // image branch (144 tokens): all queries in one 10-warp CTA, K/V streamed once
    const int smem = (160 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_fwd_kernel<10>, smem));
    attn_fwd_kernel<10><<<dim3(1, B * H), 320, smem, (cudaStream_t)stream>>>(
        (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (bf16 *)O, ldo, LSE, H, Lq, Lk, scale, drop_p, seed_ptr, op_id);
  } else if (use8(Lq)) {
    const int smem = (128 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_fwd_kernel<8>, smem));
    attn_fwd_kernel<8><<<dim3(ceil_div(Lq, 128), B * H), 256, smem, (cudaStream_t)stream>>>(
        (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (bf16 *)O, ldo, LSE, H, Lq, Lk, scale, drop_p, seed_ptr, op_id);
  } else {
    const int smem = (64 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_fwd_kernel<4>, smem));
    attn_fwd_kernel<4><<<dim3(ceil_div(Lq, 64), B * H), 128, smem, (cudaStream_t)stream>>>(
        (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (bf16 *)O, ldo, LSE, H, Lq, Lk, scale, drop_p, seed_ptr, op_id);
  }
  return check_launch("attn_fwd_kernel");
}

int vpf_attention_bwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, const void *O, int ldo,
                      const void *dO, int lddo, const float *LSE, float *delta_ws, void *dQ, int lddq, void *dK,
                      void *dV, int lddkv, int B, int H, int Lq, int Lk, int head_dim, float scale, float drop_p,
                      const unsigned long long *seed_ptr, unsigned int op_id, void *stream) {
  VPF_REQUIRE(Q && K && V && O && dO && LSE && delta_ws && dQ && dK && dV, "attention_bwd: null pointer");
  VPF_REQUIRE(head_dim == HD, "attention_bwd: head_dim=%d unsupported (64)", head_dim);
  VPF_REQUIRE((ldq % 8) == 0 && (ldkv % 8) == 0 && (ldo % 8) == 0 && (lddo % 8) == 0 && (lddq % 2) == 0 && (lddkv % 2) == 0, "attention_bwd: bad stride");
  if (B == 0) return VPF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(Lq, drop_p) && (lddq % 8) == 0 && (lddkv % 8) == 0 && aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O) &&
      aligned16(dO) && aligned16(dQ) && aligned16(dK) && aligned16(dV))
    return atc::attention_tc_bwd(Q, ldq, K, V, ldkv, O, ldo, dO, lddo, LSE, dQ, lddq, dK, dV, lddkv, B, H, Lq, Lk, scale, drop_p,
                                 seed_ptr, op_id, st);
#define FUSED_ARGS (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (const bf16 *)O, ldo, (const bf16 *)dO, lddo, LSE, (bf16 *)dQ, lddq, (bf16 *)dK, (bf16 *)dV, lddkv, H, Lq, Lk, scale, drop_p, seed_ptr, op_id
  if (Lq <= 128) {
    const int smem = (4 * 128 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_bwd_fused_kernel<8>, smem));
    attn_bwd_fused_kernel<8><<<B * H, 256, smem, st>>>(FUSED_ARGS);
    return check_launch("attn_bwd_fused_kernel<8>");
  }
  if (Lq <= 160) {   // image branch: 144 tokens
    const int smem = (4 * 160 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_bwd_fused_kernel<10>, smem));
    attn_bwd_fused_kernel<10><<<B * H, 320, smem, st>>>(FUSED_ARGS);
    return check_launch("attn_bwd_fused_kernel<10>");
  }
#undef FUSED_ARGS
  // two-kernel path (Lq > 160): delta = rowsum(dO * O) first, the fused kernel computes it itself
  const long long total = (long long)B * Lq * H;
  attn_delta_kernel<<<(unsigned)ceil_div(total, 256LL), 256, 0, st>>>((const bf16 *)O, ldo, (const bf16 *)dO, lddo, delta_ws, H, Lq, total);
  VPF_TRY(check_launch("attn_delta_kernel"));
#define DKV_ARGS (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (const bf16 *)dO, lddo, LSE, delta_ws, (bf16 *)dK, (bf16 *)dV, lddkv, H, Lq, Lk, scale, drop_p, seed_ptr, op_id
#define DQ_ARGS (const bf16 *)Q, ldq, (const bf16 *)K, (const bf16 *)V, ldkv, (const bf16 *)dO, lddo, LSE, delta_ws, (bf16 *)dQ, lddq, H, Lq, Lk, scale, drop_p, seed_ptr, op_id
  if (use8(Lk)) {
    const int smem = (2 * 128 + 4 * 64) * HD * 2 + 4 * 64 * 4;
    VPF_TRY(set_smem(attn_bwd_dkv_kernel<8>, smem));
    attn_bwd_dkv_kernel<8><<<dim3(ceil_div(Lk, 128), B * H), 256, smem, st>>>(DKV_ARGS);
  } else {
    const int smem = (2 * 64 + 4 * 64) * HD * 2 + 4 * 64 * 4;
    VPF_TRY(set_smem(attn_bwd_dkv_kernel<4>, smem));
    attn_bwd_dkv_kernel<4><<<dim3(ceil_div(Lk, 64), B * H), 128, smem, st>>>(DKV_ARGS);
  }
  VPF_TRY(check_launch("attn_bwd_dkv_kernel"));
  if (use8(Lq)) {
    const int smem = (2 * 128 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_bwd_dq_kernel<8>, smem));
    attn_bwd_dq_kernel<8><<<dim3(ceil_div(Lq, 128), B * H), 256, smem, st>>>(DQ_ARGS);
  } else {
    const int smem = (2 * 64 + 4 * 64) * HD * 2;
    VPF_TRY(set_smem(attn_bwd_dq_kernel<4>, smem));
    attn_bwd_dq_kernel<4><<<dim3(ceil_div(Lq, 64), B * H), 128, smem, st>>>(DQ_ARGS);
  }
#undef DKV_ARGS
#undef DQ_ARGS
  return check_launch("attn_bwd_dq_kernel");
}

}  // extern "C"
