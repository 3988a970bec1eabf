// attention_tc.cu -- the attention core of partseg.py:67-86 on the 5th-generation tensor cores:
//   softmax(Q K^T * scale) -> dropout(p) -> . V   per (sample, head), head dim 64, Lq <= 128 query tokens
// (the 96 / 128 latent tokens of every point-cloud layer: self-attention over 128 keys, cross-attention over
// the 1024..2500 points of the cloud), forward and backward.  tcgen05.mma issued by one elected thread, operands
// staged in shared memory by TMA (3-D maps [sample][token][channel]: out-of-range tokens are zero-filled), S / dP /
// P.V / dQ / dK / dV accumulators in TMEM, softmax arithmetic thread-per-row straight from tcgen05.ld, P / dS handed
// back to the tensor core through 128B-swizzled shared-memory tiles.  Logits never touch HBM.
//
// One shared-memory tile = [128 rows][64 bf16] (rows of 128 B, 16-byte chunks XOR-swizzled by row & 7 = the TMA
// SWIZZLE_128B pattern), used through UMMA descriptors either K-major (rows = M/N index, the 64 columns = K) or
// MN-major (rows = K index, the 64 columns = M/N), exactly the two forms gemm.cu uses:
//   forward :  S  = Q K^T      A = Q  (K-major)        B = K  (K-major)      M 128, N 128, K 64
//              O += P V        A = P  (K-major)        B = V  (MN-major)     M 128, N 64,  K 128
//   backward:  S  = Q K^T, dP = dO V^T                  (as above)
//              dV += P^T dO    A = P  (MN-major)       B = dO (MN-major)     M 128 keys, N 64, K 128 queries
//              dK += dS^T Q    A = dS (MN-major)       B = Q  (MN-major)
//              dQ += dS K      A = dS (K-major)        B = K  (MN-major)     M 128 queries, N 64, K 128 keys
// so Q, K, V, dO are loaded once and P, dS written once per (query block, key block) pair.
//
// Dropout on the attention probabilities (partseg.py:81): keep(i, j) = byte (j & 3) of hash(seed, op, (bh, i, j >> 2))
// >= round(256 p) -- one 32-bit hash serves 4 adjacent keys of a row; regenerated in the backward pass
// (same function as attention.cu; restated in oracle/rng.py).
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace vpf {
namespace atc {

typedef __nv_bfloat16 bf16;
constexpr int HD = 64;
constexpr int kTile = 128 * HD * 2;   // 16 KB
constexpr float kLog2e = 1.4426950408889634f;

// ---- small PTX helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst_s, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_s), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// producer-side wait: the TMA warp is never on the critical path (two stages of look-ahead), so it backs off instead of
// burning issue slots of the softmax warps in a try_wait spin
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  while (!ptx::mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk c (0..7) of row r inside a [128][64] bf16 tile
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

struct Drop {
  uint32_t thr, key;
  float scale;
};
__device__ __forceinline__ Drop make_drop(float p, const unsigned long long *seed_ptr, uint32_t op_id) {
  Drop d;
  d.thr = p > 0.f ? (uint32_t)(p * 256.f + 0.5f) : 0u;
  d.key = p > 0.f ? rng::make_key(seed_ptr ? *seed_ptr : 0ull, op_id) : 0u;
  d.scale = d.thr ? 256.f / (256.f - (float)d.thr) : 1.f;
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
// bit 7 of byte e set <=> byte e of the hash is < thr (element e is DROPPED); exact for thr <= 128 (SWAR, no borrows:
// every byte of kthr = 0x7f + thr is >= the 7-bit value subtracted from it)
__device__ __forceinline__ uint32_t drop_bits(uint32_t hsh, uint32_t kthr) {
  return (kthr - (hsh & 0x7f7f7f7fu)) & ~hsh & 0x80808080u;
}
// hash of the 4-key group g4 (= j >> 2) of row `rowbase` (= (bh * Lq + i) * ceil(Lk / 4))
__device__ __forceinline__ uint32_t quad_hash(const Drop &dc, uint32_t rowbase, uint32_t g4) {
  return rng::mix32((rowbase + g4) * 0x9e3779b1u ^ dc.key);
}

// UMMA issue helpers (one elected thread)
__device__ __forceinline__ void mma_kmajor_kmajor(uint32_t tmem_d, uint32_t a_s, uint32_t b_s, int ksteps, uint32_t idesc, bool acc0) {
  // A, B: K-major tiles, K advances 32 B per UMMA_K (16 bf16) inside the 128-byte swizzled row
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < ksteps)
      ptx::umma_bf16(tmem_d, ptx::umma_smem_desc(a_s + k * 32, 16, 1024), ptx::umma_smem_desc(b_s + k * 32, 16, 1024), idesc,
                     (acc0 || k > 0) ? 1u : 0u);
  }
}

// =================================================================================================== forward
// CTA = 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 softmax (thread = one query row x one 64-key half
// of the 128-key block).  Two CTAs per SM (112.6 KB shared memory, 256 TMEM columns each), persistent over
// (sample, head) items; key blocks of 128 stream through a 2-stage K/V ring.
// O stays in TMEM: the P.V products of all key blocks of an item accumulate there (tcgen05.mma accumulate), the softmax
// threads keep only the running (max, sum) of their row.  The running max is LAZY: a row switches to a larger reference
// only when the block maximum exceeds the current one by more than 2^8 (probabilities up to 256 are harmless in fp32 /
// bf16), and only then is that row of O rescaled in TMEM (tcgen05.ld -> multiply -> tcgen05.st) -- after the first key
// block of a 2048-point cross-attention this almost never happens.  Softmax arithmetic per logit: max, fma, ex2, add,
// 3 for the dropout bit, half a pack; the 1 / (1 - p) of dropout is folded into the final normalisation.
#ifdef VPF_ATTN_TIMING
// experiment-only build (VPF_NVCC_EXTRA=-DVPF_ATTN_TIMING): globaltimer stamps of CTA 0 (softmax warp 2 lane 0: slots 0..5
// per block, MMA warp lane 0: slots 6..7), first 32 blocks
__device__ unsigned long long g_attn_stamp[32][8];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ATT_STAMP(cond, n, slot) do { if ((cond) && (n) < 32) g_attn_stamp[n][slot] = gtime(); } while (0)
#else
#define ATT_STAMP(cond, n, slot) do { } while (0)
#endif

constexpr int kFwdThreads = 320;
constexpr int kFOffQ = 0, kFOffKV = kTile, kFOffP = kFOffKV + 4 * kTile, kFOffX = kFOffP + 2 * kTile, kFOffBar = kFOffX + 512;
constexpr int kFwdSmem = kFOffBar + 96;
static_assert(kFwdSmem <= 115712, "two forward CTAs per SM");
constexpr float kLazy = 8.f;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

// rare path of the lazy running max: multiply this thread's 32 O columns (one TMEM lane) by corr.  Not inlined, so its
// 32 registers do not add to the pressure of the softmax loop.
__device__ __noinline__ void rescale_o(uint32_t taddr, float corr) {
  uint32_t o[32];
  ptx::tmem_ld_32x32(taddr, o);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * corr);
  tmem_st32(taddr, o);
  tmem_st_wait();
}

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, bf16 *__restrict__ O, int ldo, float *__restrict__ LSE,
                   int H, int Lq, int Lk, int nqb, int n_items, float scale, float drop_p,
                   const unsigned long long *__restrict__ seed_ptr, uint32_t op_id) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_s = ptx::smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kFOffBar);
  uint64_t *q_full = bars + 0, *q_empty = bars + 1, *kv_full = bars + 2, *kv_empty = bars + 4, *s_full = bars + 6,
           *s_empty = bars + 7, *p_full = bars + 8, *pv_full = bars + 9, *o_empty = bars + 10;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 11);
  __nv_bfloat16 *xm = reinterpret_cast<__nv_bfloat16 *>(smem + kFOffX);   // [2][128] block-max bounds of the two halves
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (Lk + 127) >> 7;

  if (threadIdx.x == 0 && (smem_s & 1023u)) __trap();   // the swizzled tiles need a 1024-byte aligned base
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      ptx::mbar_init(q_full, 1); ptx::mbar_init(q_empty, 1);
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
      ptx::mbar_init(s_full, 1); ptx::mbar_init(s_empty, 8); ptx::mbar_init(p_full, 8);
      ptx::mbar_init(pv_full, 1); ptx::mbar_init(o_empty, 8);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_s + kFOffQ, sKV = smem_s + kFOffKV, sP = smem_s + kFOffP;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int n = 0, j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      const int bh = it / nqb, qb = it - bh * nqb, b = bh / H, h = bh - b * H;   // item = (sample, head, query block)
      mbar_wait_backoff(q_empty, (j & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(q_full, kTile);
        tma_load_3d(sQ, &tmQ, q_full, h * HD, qb * 128, b);
      }
      __syncwarp();
      for (int t = 0; t < nkb; ++t, ++n) {
        const int st = n & 1;
        mbar_wait_backoff(&kv_empty[st], ((n >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
          tma_load_3d(sKV + st * 2 * kTile, &tmK, &kv_full[st], h * HD, t * 128, b);
          tma_load_3d(sKV + st * 2 * kTile + kTile, &tmV, &kv_full[st], h * HD, t * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 64, 0, 1);
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_items * nkb;
    auto issue_s = [&](int n) {
      const int t = n % nkb, j = n / nkb, st = n & 1;
      const int kv16 = (min(128, Lk - t * 128) + 15) & ~15;     // valid keys of this block, rounded up to the UMMA N step
      if (t == 0) ptx::mbar_wait(q_full, j & 1);
      ptx::mbar_wait(&kv_full[st], (n >> 1) & 1);
      ptx::mbar_wait(s_empty, (n & 1) ^ 1);
      ptx::tc_fence_after();
      ATT_STAMP(blockIdx.x == 0 && lane == 0, n, 6);
      if (ptx::elect_one()) {
        mma_kmajor_kmajor(tmem_base, sQ, sKV + st * 2 * kTile, 4, ptx::umma_idesc_bf16(128, kv16, 0, 0), false);
        ptx::umma_commit(s_full);
        if (t == nkb - 1) ptx::umma_commit(q_empty);
      }
      __syncwarp();
    };
    if (total > 0) issue_s(0);
    for (int n = 0; n < total; ++n) {
      if (n + 1 < total) issue_s(n + 1);
      const int t = n % nkb, j = n / nkb, st = n & 1;
      const int ksteps = (min(128, Lk - t * 128) + 15) >> 4;
      ptx::mbar_wait(p_full, n & 1);
      if (t == 0) ptx::mbar_wait(o_empty, (j & 1) ^ 1);      // the previous item's O has been read out of TMEM
      ptx::tc_fence_after();
      ATT_STAMP(blockIdx.x == 0 && lane == 0, n, 7);
      if (ptx::elect_one()) {
        const uint32_t sV = sKV + st * 2 * kTile + kTile;
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 128 keys = 8 UMMA_K steps: P k-block (k >> 2), V rows 16 k .. 16 k + 15
          if (k < ksteps) {
            const uint64_t ad = ptx::umma_smem_desc(sP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t bd = ptx::umma_smem_desc(sV + k * 2048, kTile, 1024);
            ptx::umma_bf16(tmem_base + 128, ad, bd, idesc_pv, (t > 0 || k > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(pv_full);
        ptx::umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 2, quad = warp & 3, hf = sw >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const Drop dc = make_drop(drop_p, seed_ptr, op_id);
    const float sc2 = scale * kLog2e;
    const uint32_t kq4 = (uint32_t)((Lk + 3) >> 2);
    const uint32_t prow = sP + hf * kTile;
    const uint32_t kthr = 0x7f7f7f7fu + dc.thr * 0x01010101u;
    // The end of an item (wait for its last P.V, read O out of TMEM, normalise, store) is DEFERRED into the first block of
    // the next item, behind that block's logit load / maximum exchange: the tensor-core latency of the last P.V and the
    // skew between the eight softmax warps are hidden there instead of idling this CTA.  The two threads of a row
    // exchange their partial row sums through two spare TMEM columns (192 + half), ordered by the p_full -> pv_full chain.
    struct { bool on; int n, b, h, bh, qrow; float m, l; } pend = {false, 0, 0, 0, 0, 0, 0.f, 0.f};
    auto epilogue = [&]() {
      ptx::mbar_wait(pv_full, pend.n & 1);
      ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, pend.n, 4);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld_32x32(t_row + 128 + hf * 32, o);
      const float l_other = __uint_as_float(tmem_ld1(t_row + 192 + (hf ^ 1)));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);
      if (pend.qrow < Lq) {
        const float tot = pend.l + l_other;
        const float inv = dc.scale / tot;
        bf16 *op = O + ((size_t)pend.b * Lq + pend.qrow) * ldo + pend.h * HD + hf * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          v.y = pack2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          v.z = pack2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          v.w = pack2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          *reinterpret_cast<uint4 *>(op + 8 * c) = v;
        }
        if (hf == 0) LSE[(size_t)pend.bh * Lq + pend.qrow] = pend.m + log2f(tot);
      }
      ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, pend.n, 5);
      pend.on = false;
    };
    int n = 0, j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      const int bh = it / nqb, qb = it - bh * nqb, b = bh / H, h = bh - b * H;
      const int qrow = qb * 128 + row;                           // this thread's query token
      const bool active = qb * 128 + quad * 32 < Lq;             // warp-uniform: any valid query row in this warp?
      const uint32_t rowbase = ((uint32_t)bh * (uint32_t)Lq + (uint32_t)qrow) * kq4;
      float m = 0.f, l = 0.f;
      for (int t = 0; t < nkb; ++t, ++n) {
        const int kvalid = min(128, Lk - t * 128);
        const int j0 = hf * 64;                         // this thread's key offset inside the block
        ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, n, 0);
        ptx::mbar_wait(s_full, n & 1);
        ptx::tc_fence_after();
        ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, n, 1);
        if (!active) {
          // no valid query row in this warp (tail query block): keep the pipeline protocol, skip the arithmetic; the
          // P rows stay whatever they were -- row m of P only reaches row m of O, which is never stored
          named_bar(1 + quad, 64);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) { ptx::mbar_arrive(s_empty); }
          if (pend.on) epilogue();
          else if (n > 0) ptx::mbar_wait(pv_full, (n - 1) & 1);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(p_full);
          continue;
        }
        // block maximum of this half (raw logits; sc2 > 0).  Only 32 logits are kept in registers at a time: the upper
        // 32 of this thread's 64 keys are read twice (maximum now, probabilities later) -- a tcgen05.ld is cheaper than
        // the spills 64 live logits cost under the 96-register budget of two CTAs per SM.
        uint32_t sx[32];
        float mx = -INFINITY;
        ptx::tmem_ld_32x32(t_row + j0 + 32, sx);
        ptx::tmem_ld_wait();
        if (kvalid == 128) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(sx[e]), __uint_as_float(sx[e + 1])));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) if (j0 + 32 + e < kvalid) mx = fmaxf(mx, __uint_as_float(sx[e]));
        }
        ptx::tmem_ld_32x32(t_row + j0, sx);
        ptx::tmem_ld_wait();
        if (kvalid == 128) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(sx[e]), __uint_as_float(sx[e + 1])));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) if (j0 + e < kvalid) mx = fmaxf(mx, __uint_as_float(sx[e]));
        }
        mx *= sc2;
        // exchanged with the other half of the row as a bf16 UPPER bound
        const float bound = mx + fabsf(mx) * 0.0078125f;          // >= mx after round-to-nearest to 8 mantissa bits
        xm[hf * 128 + row] = __float2bfloat16(mx == -INFINITY ? -1e30f : bound);
        named_bar(1 + quad, 64);
        const float m_blk = fmaxf(__bfloat162float(xm[row]), __bfloat162float(xm[128 + row]));
        ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, n, 2);
        bool need = t == 0 ? false : (m_blk > m + kLazy);
        if (t == 0) m = m_blk;
        if (__any_sync(0xffffffffu, need)) {
          // rescale this row of O in TMEM (the previous P.V must have retired) and the running sum
          const float m_new = need ? m_blk : m;
          const float corr = ex2(m - m_new);
          ptx::mbar_wait(pv_full, (n - 1) & 1);
          ptx::tc_fence_after();
          rescale_o(t_row + 128 + hf * 32, corr);
          l *= corr;
          m = m_new;
        }
        // the previous item leaves TMEM here (its last P.V has had the whole load / exchange phase to finish); otherwise
        // just make sure the previous P.V of this item has finished reading the P tile
        if (pend.on) epilogue();
        else if (n > 0) ptx::mbar_wait(pv_full, (n - 1) & 1);
        // probabilities -> dropout -> bf16 P tile
        float lsum = 0.f;
        const float negm = -m;
        const uint32_t g4 = (uint32_t)((t * 128 + j0) >> 2);
        auto half = [&](const uint32_t (&sx)[32], int hh, bool tail) {   // 32 keys: 4 chunks of 8
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              p[e] = ex2(fmaf(__uint_as_float(sx[c * 8 + e]), sc2, negm));
              if (tail && j0 + hh * 32 + c * 8 + e >= kvalid) p[e] = 0.f;
              lsum += p[e];
            }
            uint32_t w0 = pack2(p[0], p[1]), w1 = pack2(p[2], p[3]), w2 = pack2(p[4], p[5]), w3 = pack2(p[6], p[7]);
            if (dc.thr) {
              const uint32_t d0 = drop_bits(quad_hash(dc, rowbase, g4 + hh * 8 + c * 2), kthr);
              const uint32_t d1 = drop_bits(quad_hash(dc, rowbase, g4 + hh * 8 + c * 2 + 1), kthr);
              w0 &= ~prmt(d0, 0x9988u); w1 &= ~prmt(d0, 0xbbaau);
              w2 &= ~prmt(d1, 0x9988u); w3 &= ~prmt(d1, 0xbbaau);
            }
            sts128(prow + tile_off(row, hh * 4 + c), w0, w1, w2, w3);
          }
        };
        if (kvalid == 128) half(sx, 0, false); else half(sx, 0, true);
        ptx::tmem_ld_32x32(t_row + j0 + 32, sx);
        ptx::tmem_ld_wait();
        // S has been read for the last time: the next block's logits may be issued.  (This arrive comes after the xm
        // exchange read above, which also orders the next block's xm writes -- they need the next s_full -- behind it.)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_empty);
        if (kvalid == 128) half(sx, 1, false); else half(sx, 1, true);
        l += lsum;
        if (t == nkb - 1) {      // this half's row sum for the other half's thread (read after this item's last pv_full)
          tmem_st1(t_row + 192 + hf, __float_as_uint(l));
          tmem_st_wait();
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(p_full);
        ATT_STAMP(blockIdx.x == 0 && warp == 2 && lane == 0, n, 3);
      }
      pend.on = true; pend.n = n - 1; pend.b = b; pend.h = h; pend.bh = bh; pend.qrow = qrow; pend.m = m; pend.l = l;
    }
    if (pend.on) epilogue();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// =================================================================================================== backward
// CTA = 18 warps, one per SM, persistent over (sample, head) items: warp 0 TMA, warp 1 MMA, warps 2..17 compute
// (thread = one query row x 32 of the 128 keys of the block).  TMEM: S 0..127, dP 128..255, dV 256..319, dK 320..383,
// dQ 384..447.  Per (item, key block):
//   MMA      S = Q K^T, dP = dO V^T                                         (issued one block ahead)
//   compute  P = exp2(S sc - lse), dS = P (drop(dP) - delta) scale  ->  bf16 P / dS tiles in shared memory
//   MMA      dV = P^T dO, dK = dS^T Q   (flushed every key block: all Lq <= 128 queries are in this block)
//            dQ += dS K                 (accumulates over the key blocks of the item)
//   compute  flush of block n-1 (TMEM -> registers -> global) rides behind the tile stores of block n, so the tensor
//            pipe always has the next group queued.
// Q, dO and O arrive together by TMA (delta = rowsum(dO o O) is computed from the shared-memory tiles); the 1 / (1 - p)
// of dropout is folded into delta, the logit scale and the dV flush.
constexpr int kBwdThreads = 576;
constexpr int kBOffQdO = 0, kBOffKV = 6 * kTile, kBOffP = 10 * kTile, kBOffdS = 12 * kTile, kBOffBar = 14 * kTile;
constexpr int kBwdSmem = kBOffBar + 256;
static_assert(kBwdSmem <= 232448, "backward shared memory");

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmO, const float *__restrict__ LSE, bf16 *__restrict__ dQ, int lddq,
                   bf16 *__restrict__ dK, bf16 *__restrict__ dV, int lddkv, int H, int Lq, int Lk, int nqb, int n_items,
                   float scale, float drop_p, const unsigned long long *__restrict__ seed_ptr, uint32_t op_id) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_s = ptx::smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kBOffBar);
  uint64_t *qdo_full = bars + 0, *qdo_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 6, *sdp_full = bars + 8,
           *sdp_empty = bars + 9, *pds_full = bars + 10, *acc_full = bars + 11, *acc_empty = bars + 12;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 13);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (Lk + 127) >> 7;

  if (threadIdx.x == 0 && (smem_s & 1023u)) __trap();
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmdO); ptx::prefetch_tmap(&tmO);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&qdo_full[s], 1); ptx::mbar_init(&qdo_empty[s], 1);
        ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1);
      }
      ptx::mbar_init(sdp_full, 1); ptx::mbar_init(sdp_empty, 16); ptx::mbar_init(pds_full, 16);
      ptx::mbar_init(acc_full, 1); ptx::mbar_init(acc_empty, 16);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQdO = smem_s + kBOffQdO, sKV = smem_s + kBOffKV, sP = smem_s + kBOffP, sdS = smem_s + kBOffdS;
  // One item = one (sample, head).  Its steps walk the key blocks (outer) and the query blocks (inner, nqb <= 2):
  // dV / dK accumulate over the query blocks of a key block and are flushed after the last one; dQ of query block qb
  // lives in its own TMEM columns (384 + 64 qb), accumulates over the key blocks and is flushed after the last one.
  const int spi = nkb * nqb;   // steps per item
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_items > 0 ? my_items * spi : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      const int b = it / H, h = it - b * H;
      // loads are issued in the order the steps consume them (query block qb at its first key block, key block t at its
      // first query block): the MMA warp looks one step ahead, so the first step of the next item must never wait behind
      // a buffer that only the current item's LAST step releases
      for (int u = 0; u < spi; ++u) {
        const int t = u / nqb, qb = u - t * nqb;
        if (t == 0) {
          const int qi = j * nqb + qb, qs = qi & 1;
          mbar_wait_backoff(&qdo_empty[qs], ((qi >> 1) & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&qdo_full[qs], 3 * kTile);
            tma_load_3d(sQdO + qs * 3 * kTile, &tmQ, &qdo_full[qs], h * HD, qb * 128, b);
            tma_load_3d(sQdO + qs * 3 * kTile + kTile, &tmdO, &qdo_full[qs], h * HD, qb * 128, b);
            tma_load_3d(sQdO + qs * 3 * kTile + 2 * kTile, &tmO, &qdo_full[qs], h * HD, qb * 128, b);
          }
          __syncwarp();
        }
        if (qb == 0) {
          const int c = j * nkb + t, st = c & 1;
          mbar_wait_backoff(&kv_empty[st], ((c >> 1) & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
            tma_load_3d(sKV + st * 2 * kTile, &tmK, &kv_full[st], h * HD, t * 128, b);
            tma_load_3d(sKV + st * 2 * kTile + kTile, &tmV, &kv_full[st], h * HD, t * 128, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_t = ptx::umma_idesc_bf16(128, 64, 1, 1);    // A = P / dS transposed (MN-major), B MN-major
    const uint32_t idesc_q = ptx::umma_idesc_bf16(128, 64, 0, 1);    // A = dS K-major, B = K MN-major
    auto issue_sdp = [&](int n) {
      const int j = n / spi, u = n - j * spi, t = u / nqb, qb = u - t * nqb;
      const int qi = j * nqb + qb, qs = qi & 1, c = j * nkb + t, st = c & 1;
      const int kv16 = (min(128, Lk - t * 128) + 15) & ~15;
      if (t == 0) ptx::mbar_wait(&qdo_full[qs], (qi >> 1) & 1);
      if (qb == 0) ptx::mbar_wait(&kv_full[st], (c >> 1) & 1);
      ptx::mbar_wait(sdp_empty, (n & 1) ^ 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t q_s = sQdO + qs * 3 * kTile, do_s = q_s + kTile, k_s = sKV + st * 2 * kTile, v_s = k_s + kTile;
        const uint32_t idesc_s = ptx::umma_idesc_bf16(128, kv16, 0, 0);
        mma_kmajor_kmajor(tmem_base, q_s, k_s, 4, idesc_s, false);          // S  = Q K^T
        mma_kmajor_kmajor(tmem_base + 128, do_s, v_s, 4, idesc_s, false);   // dP = dO V^T
        ptx::umma_commit(sdp_full);
      }
      __syncwarp();
    };
    if (total > 0) issue_sdp(0);
    for (int n = 0; n < total; ++n) {
      if (n + 1 < total) issue_sdp(n + 1);
      const int j = n / spi, u = n - j * spi, t = u / nqb, qb = u - t * nqb;
      const int qi = j * nqb + qb, qs = qi & 1, c = j * nkb + t, st = c & 1;
      const int ksteps = (min(128, Lk - t * 128) + 15) >> 4;
      ptx::mbar_wait(pds_full, n & 1);
      ptx::mbar_wait(acc_empty, (n & 1) ^ 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t q_s = sQdO + qs * 3 * kTile, do_s = q_s + kTile, k_s = sKV + st * 2 * kTile;
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // K = 128 queries: MN-major A advances 16 rows = 2048 B per step
          const uint64_t ap = ptx::umma_smem_desc(sP + k * 2048, kTile, 1024);
          const uint64_t as = ptx::umma_smem_desc(sdS + k * 2048, kTile, 1024);
          const uint64_t bo = ptx::umma_smem_desc(do_s + k * 2048, kTile, 1024);
          const uint64_t bq = ptx::umma_smem_desc(q_s + k * 2048, kTile, 1024);
          ptx::umma_bf16(tmem_base + 256, ap, bo, idesc_t, (qb > 0 || k > 0) ? 1u : 0u);   // dV += P^T dO
          ptx::umma_bf16(tmem_base + 320, as, bq, idesc_t, (qb > 0 || k > 0) ? 1u : 0u);   // dK += dS^T Q
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // K = the block's keys: dS K-major k-block (k >> 2), K rows 16 k ..
          if (k < ksteps) {
            const uint64_t ad = ptx::umma_smem_desc(sdS + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t bk = ptx::umma_smem_desc(k_s + k * 2048, kTile, 1024);
            ptx::umma_bf16(tmem_base + 384 + qb * 64, ad, bk, idesc_q, (t > 0 || k > 0) ? 1u : 0u);   // dQ[qb] += dS K
          }
        }
        ptx::umma_commit(acc_full);
        if (qb == nqb - 1) ptx::umma_commit(&kv_empty[st]);
        if (t == nkb - 1) ptx::umma_commit(&qdo_empty[qs]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ compute warps
    const int cw = warp - 2, quad = warp & 3, qt = cw >> 2;     // qt: which 32 of the block's 128 keys
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const Drop dc = make_drop(drop_p, seed_ptr, op_id);
    const float sc2 = scale * kLog2e;
    const float scale_d = scale * dc.scale, inv_dscale = 1.f / dc.scale;
    const uint32_t kq4 = (uint32_t)((Lk + 3) >> 2);
    const uint32_t kthr = 0x7f7f7f7fu + dc.thr * 0x01010101u;
    auto pack8 = [](const uint32_t *r, float f) {
      uint4 a;
      a.x = pack2(__uint_as_float(r[0]) * f, __uint_as_float(r[1]) * f); a.y = pack2(__uint_as_float(r[2]) * f, __uint_as_float(r[3]) * f);
      a.z = pack2(__uint_as_float(r[4]) * f, __uint_as_float(r[5]) * f); a.w = pack2(__uint_as_float(r[6]) * f, __uint_as_float(r[7]) * f);
      return a;
    };
    // flush of one finished block: dV, dK (rows = keys of that block) and, at the end of an item, dQ (rows = queries)
    auto flush = [&](int fn) {
      const int fj = fn / spi, fu = fn - fj * spi, ft = fu / nqb, fq = fu - ft * nqb;
      const int fit = (int)blockIdx.x + fj * (int)gridDim.x;
      const int fb = fit / H, fh = fit - fb * H;
      const bool do_kv = fq == nqb - 1, do_q = ft == nkb - 1;      // warp-uniform
      uint32_t rv[16], rk[16], rq[16];
      tmem_ld16(t_row + 256 + qt * 16, rv);      // (unconditional loads keep the three arrays in registers)
      tmem_ld16(t_row + 320 + qt * 16, rk);
      tmem_ld16(t_row + 384 + fq * 64 + qt * 16, rq);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
      const int key = ft * 128 + row, qrow = fq * 128 + row;
      if (do_kv && key < Lk) {
        uint4 *pk = reinterpret_cast<uint4 *>(dK + ((size_t)fb * Lk + key) * lddkv + fh * HD + qt * 16);
        uint4 *pv = reinterpret_cast<uint4 *>(dV + ((size_t)fb * Lk + key) * lddkv + fh * HD + qt * 16);
        pk[0] = pack8(rk, scale_d); pk[1] = pack8(rk + 8, scale_d);
        pv[0] = pack8(rv, dc.scale); pv[1] = pack8(rv + 8, dc.scale);
      }
      if (do_q && qrow < Lq) {
        uint4 *pq = reinterpret_cast<uint4 *>(dQ + ((size_t)fb * Lq + qrow) * lddq + fh * HD + qt * 16);
        pq[0] = pack8(rq, scale_d); pq[1] = pack8(rq + 8, scale_d);
      }
    };
    int n = 0, j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      // per-row constants of the item's query blocks: lse (log2 units) and delta' = sum_d dO[i, d] O[i, d] / dropout scale
      float lse = INFINITY, delta = 0.f;     // rows >= Lq: p = exp2(-inf) = 0, dS = 0
      for (int u = 0; u < spi; ++u, ++n) {
        const int t = u / nqb, qb = u - t * nqb;
        const int qi = j * nqb + qb, qs = qi & 1;
        const int qrow = qb * 128 + row;
        const uint32_t rowbase = ((uint32_t)it * (uint32_t)Lq + (uint32_t)qrow) * kq4;
        if (t == 0 || nqb > 1) {     // (one query block: once per item; two: the blocks alternate, recompute per step)
          lse = INFINITY; delta = 0.f;
          if (qrow < Lq) lse = __ldg(LSE + (size_t)it * Lq + qrow);
          if (t == 0) ptx::mbar_wait(&qdo_full[qs], (qi >> 1) & 1);
          const uint32_t do_s = sQdO + qs * 3 * kTile + kTile, o_s = do_s + kTile;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 a = lds128(o_s + tile_off(row, c)), d = lds128(do_s + tile_off(row, c));
            const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&a), *hd = reinterpret_cast<const __nv_bfloat162 *>(&d);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 fa = __bfloat1622float2(ha[q]), fd = __bfloat1622float2(hd[q]);
              delta = fmaf(fa.x, fd.x, delta);
              delta = fmaf(fa.y, fd.y, delta);
            }
          }
          delta *= inv_dscale;
        }
        const int kvalid = min(128, Lk - t * 128);
        // warp-uniform: no valid query row in this warp (tail query block) or no valid key in this quarter (tail key block)
        const bool dead = (qb * 128 + quad * 32 >= Lq) || (qt * 32 >= kvalid);
        const uint32_t kb = (uint32_t)(qt >> 1) * kTile, c0 = (qt & 1) * 4;
        ptx::mbar_wait(sdp_full, n & 1);
        ptx::tc_fence_after();
        if (dead) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(sdp_empty);
          if (n > 0) { ptx::mbar_wait(acc_full, (n - 1) & 1); ptx::tc_fence_after(); }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts128(sP + kb + tile_off(row, c0 + c), 0u, 0u, 0u, 0u);
            sts128(sdS + kb + tile_off(row, c0 + c), 0u, 0u, 0u, 0u);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(pds_full);
          if (n > 0) flush(n - 1);
          continue;
        }
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32(t_row + qt * 32, rs);
        ptx::tmem_ld_32x32(t_row + 128 + qt * 32, rp);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(sdp_empty);     // S / dP are in registers: the next block's may be issued
        const float neglse = -lse;
        uint32_t pp[16], pd[16];
        const uint32_t g4 = (uint32_t)((t * 128 + qt * 32) >> 2);
        auto body = [&](bool tail) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float p[4], ds[4];
            uint32_t db = 0;
            if (dc.thr) db = drop_bits(quad_hash(dc, rowbase, g4 + g), kthr);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              p[e] = ex2(fmaf(__uint_as_float(rs[g * 4 + e]), sc2, neglse));
              uint32_t dpb = rp[g * 4 + e];
              if (tail && qt * 32 + g * 4 + e >= kvalid) { p[e] = 0.f; dpb = 0u; }   // (stale TMEM columns beyond the block's keys)
              if (dc.thr) dpb &= ~prmt(db, 0x8888u + 0x1111u * e);     // dropped element: dP = 0
              ds[e] = p[e] * (__uint_as_float(dpb) - delta);            // x scale / (1 - p) at the dQ / dK flush
            }
            uint32_t w0 = pack2(p[0], p[1]), w1 = pack2(p[2], p[3]);
            if (dc.thr) { w0 &= ~prmt(db, 0x9988u); w1 &= ~prmt(db, 0xbbaau); }   // kept probabilities (x 1/(1-p) at the dV flush)
            pp[2 * g] = w0; pp[2 * g + 1] = w1;
            pd[2 * g] = pack2(ds[0], ds[1]); pd[2 * g + 1] = pack2(ds[2], ds[3]);
          }
        };
        if (kvalid == 128) body(false); else body(true);
        // the P / dS tiles are free once the MMA group of block n-1 has retired
        if (n > 0) { ptx::mbar_wait(acc_full, (n - 1) & 1); ptx::tc_fence_after(); }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          sts128(sP + kb + tile_off(row, c0 + c), pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
          sts128(sdS + kb + tile_off(row, c0 + c), pd[4 * c], pd[4 * c + 1], pd[4 * c + 2], pd[4 * c + 3]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(pds_full);
        if (n > 0) flush(n - 1);      // ... and that block's accumulators leave TMEM before the next group overwrites them
      }
    }
    if (total > 0) {
      ptx::mbar_wait(acc_full, (total - 1) & 1);
      ptx::tc_fence_after();
      flush(total - 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [B][L][H*64] bf16 view with row stride ld elements; box = 64 channels x 128 tokens x 1 sample, 128B swizzle
static int make_map3(CUtensorMap *m, const void *base, int B, int L, int H, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 7))
    return fail(VPF_EINVAL, "attention operand must be 16-byte aligned with a row stride that is a multiple of 8 (ld=%d)", ld);
  cuuint64_t dims[3] = {(cuuint64_t)H * HD, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
  cuuint32_t box[3] = {HD, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled (attention) failed (%d) B=%d L=%d H=%d ld=%d", (int)r, B, L, H, ld);
  return VPF_OK;
}

int attention_tc_fwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, void *O, int ldo, float *LSE, int B,
                     int H, int Lq, int Lk, float scale, float drop_p, const unsigned long long *seed_ptr,
                     unsigned int op_id, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  VPF_TRY(make_map3(&tq, Q, B, Lq, H, ldq));
  VPF_TRY(make_map3(&tk, K, B, Lk, H, ldkv));
  VPF_TRY(make_map3(&tv, V, B, Lk, H, ldkv));
  static bool attr = false;
  if (!attr) {
    VPF_CUDA_TRY(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr = true;
  }
  const int nqb = (Lq + 127) / 128;
  const int items = B * H * nqb;
  const int grid = min(items, 2 * num_sms());
  attn_tc_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, st>>>(tq, tk, tv, (bf16 *)O, ldo, LSE, H, Lq, Lk, nqb, items, scale, drop_p, seed_ptr, op_id);
  return check_launch("attn_tc_fwd_kernel");
}

int attention_tc_bwd(const void *Q, int ldq, const void *K, const void *V, int ldkv, const void *O, int ldo, const void *dO,
                     int lddo, const float *LSE, void *dQ, int lddq, void *dK, void *dV, int lddkv, int B, int H, int Lq,
                     int Lk, float scale, float drop_p, const unsigned long long *seed_ptr, unsigned int op_id,
                     cudaStream_t st) {
  CUtensorMap tq, tk, tv, tdo, to;
  VPF_TRY(make_map3(&tq, Q, B, Lq, H, ldq));
  VPF_TRY(make_map3(&tk, K, B, Lk, H, ldkv));
  VPF_TRY(make_map3(&tv, V, B, Lk, H, ldkv));
  VPF_TRY(make_map3(&tdo, dO, B, Lq, H, lddo));
  VPF_TRY(make_map3(&to, O, B, Lq, H, ldo));
  VPF_REQUIRE((lddq & 7) == 0 && (lddkv & 7) == 0 && (ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(dQ) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(dK) & 15) == 0 && (reinterpret_cast<uintptr_t>(dV) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(O) & 15) == 0, "attention_bwd (tcgen05): outputs must be 16-byte aligned, strides multiples of 8");
  static bool attr = false;
  if (!attr) {
    VPF_CUDA_TRY(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr = true;
  }
  const int items = B * H, nqb = (Lq + 127) / 128;
  VPF_REQUIRE(nqb <= 2, "attention_bwd (tcgen05): at most 256 query tokens (two dQ accumulators in TMEM), got %d", Lq);
  const int grid = min(items, num_sms());
  attn_tc_bwd_kernel<<<grid, kBwdThreads, kBwdSmem, st>>>(tq, tk, tv, tdo, to, LSE, (bf16 *)dQ, lddq, (bf16 *)dK, (bf16 *)dV, lddkv, H,
                                                           Lq, Lk, nqb, items, scale, drop_p, seed_ptr, op_id);
  return check_launch("attn_tc_bwd_kernel");
}

}  // namespace atc
}  // namespace vpf

#ifdef VPF_ATTN_TIMING
extern "C" int vpf_debug_attn_stamps(unsigned long long *out256) {
  VPF_CUDA_TRY(cudaDeviceSynchronize());
  VPF_CUDA_TRY(cudaMemcpyFromSymbol(out256, vpf::atc::g_attn_stamp, sizeof(unsigned long long) * 256));
  return VPF_OK;
}
#endif
