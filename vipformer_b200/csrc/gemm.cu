// gemm.cu -- the dense-contraction workhorse: bf16 x bf16 -> fp32 on tcgen05.
//
//   C[M,N] (op)= epilogue( alpha * sum_k A(m,k) * B(n,k) )
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 5-stage ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, M=128 N=128 K=16; accumulators in TMEM,
//                               two 128-column accumulator stages so the epilogue of tile i
//                               overlaps the main loop of tile i+1)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b -> registers -> smem transpose -> fused epilogue -> coalesced global)
//
// Both operands may be K-major (row-major [rows][K]) or MN-major ([K][rows]); that covers the
// forward (X.W^T), the data gradient (dY.W) and the weight gradient (dY^T.X, split-K with
// red.global.add.v4.f32) of every Linear / 1x1-Conv on the ViPFormer hot path
// (vipformer/model/pointcloud/partseg.py:15-198, utils.py:144-189, classifier.py:25-50)
// without materialising a single transpose.
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace vpf {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 5;
constexpr int kTileBytesA = BM * BK * 2, kTileBytesB = BN * BK * 2;
constexpr int kStageBytes = kTileBytesA + kTileBytesB;
constexpr int kGemmThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kStgLd = 36;   // padded row stride (floats) of the per-warp epilogue transpose tile
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 32 * kStgLd * 4 /*epilogue staging*/;
constexpr int kTmemCols = 2 * BN;

struct GemmArgs {
  int M, N, K;
  int a_mn, b_mn;
  int num_m_tiles, num_n_tiles, kblocks, kblocks_per_split, splits;
  vpf_gemm_epilogue e;
};

// exact-erf GELU (nn.GELU default, partseg.py:196) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below
// the bf16 rounding of the stored result); one MUFU.EX2 + one MUFU.RCP instead of the ~30-instruction erff().
__device__ __forceinline__ void erf_parts(float x, float &erf_v, float &gauss) {
  // for z = x / sqrt(2): erf(z) and exp(-z^2) = exp(-x^2 / 2)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  gauss = __expf(-z * z);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * gauss;
  erf_v = copysignf(e, x);
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

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *tmem_full = empty_bar + kStages;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 8); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = g.num_m_tiles * g.num_n_tiles * g.splits;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int n_tile = w % g.num_n_tiles;
      const int t = w / g.num_n_tiles;
      const int m_tile = t % g.num_m_tiles;
      const int split = t / g.num_m_tiles;
      const int kb0 = split * g.kblocks_per_split, kb1 = min(kb0 + g.kblocks_per_split, g.kblocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          uint8_t *sa = smem + stage * kStageBytes, *sb = sa + kTileBytesA;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          if (!g.a_mn) {
            ptx::tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_tile * BM);
          } else {
            ptx::tma_load_2d(sa, &tma_a, &full_bar[stage], m_tile * BM, kb * BK);
            ptx::tma_load_2d(sa + kTileBytesA / 2, &tma_a, &full_bar[stage], m_tile * BM + 64, kb * BK);
          }
          if (!g.b_mn) {
            ptx::tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n_tile * BN);
          } else {
            ptx::tma_load_2d(sb, &tma_b, &full_bar[stage], n_tile * BN, kb * BK);
            ptx::tma_load_2d(sb + kTileBytesB / 2, &tma_b, &full_bar[stage], n_tile * BN + 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    const uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, g.a_mn, g.b_mn);
    // K-major tile: rows of 128 B, 8-row swizzle atoms 1024 B apart; one UMMA_K (16 bf16) = +32 B.
    // MN-major tile: two 64-wide blocks 8 KB apart (LBO), 8-k-row groups 1024 B apart (SBO);
    //                one UMMA_K = 16 k-rows = +2048 B.
    const uint32_t a_lbo = g.a_mn ? kTileBytesA / 2 : 16, b_lbo = g.b_mn ? kTileBytesB / 2 : 16;
    const uint32_t a_adv = g.a_mn ? 2048 : 32, b_adv = g.b_mn ? 2048 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++iter) {
      const int split = (w / g.num_n_tiles) / g.num_m_tiles;
      const int kb0 = split * g.kblocks_per_split, kb1 = min(kb0 + g.kblocks_per_split, g.kblocks);
      const int acc = iter & 1;
      ptx::mbar_wait(&tmem_empty[acc], ((iter >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes), sb = sa + kTileBytesA;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = ptx::umma_smem_desc(sa + k * a_adv, a_lbo, 1024);
            const uint64_t bdesc = ptx::umma_smem_desc(sb + k * b_adv, b_lbo, 1024);
            ptx::umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);                    // smem slot free once these MMAs retire
          if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    // TMEM gives each lane one ROW (32 fp32 columns per tcgen05.ld).  Global traffic wants lanes along a row,
    // so every 32x32 chunk is transposed through a padded per-warp shared-memory tile: afterwards 8 lanes cover
    // 128 contiguous bytes of one output row and a warp instruction touches 4 rows -> fully coalesced float4 /
    // bf16x4 loads (bias, residual, aux) and stores.  The per-patch max pool reads the same tile column-wise.
    const vpf_gemm_epilogue &e = g.e;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    float *stg = reinterpret_cast<float *>(smem + kStages * kStageBytes + 256) + (warp - 2) * (32 * kStgLd);
    uint32_t drop_thr = 0, drop_key = 0;
    float drop_scale = 1.f;
    if (e.mode == VPF_EPI_RESIDUAL && e.drop_p > 0.f) {
      drop_thr = rng::threshold(e.drop_p);
      drop_key = rng::make_key(e.seed_ptr ? *e.seed_ptr : 0ull, e.op_id);
      drop_scale = 1.f / (1.f - e.drop_p);
    }
    const int cq = (lane & 7) * 4, rsub = lane >> 3;
    const bool vec_ok = (e.ldc & 3) == 0;
    int iter = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++iter) {
      const int n_tile = w % g.num_n_tiles;
      const int m_tile = (w / g.num_n_tiles) % g.num_m_tiles;
      const int acc = iter & 1;
      ptx::mbar_wait(&tmem_full[acc], (iter >> 1) & 1);
      ptx::tc_fence_after();
      const long long row0 = (long long)m_tile * BM + quad * 32;   // first row of this warp's 32-row slab
      const int c_begin = ((warp - 2) >> 2) * (BN / 2);   // warps 2..5 take columns 0..63, warps 6..9 columns 64..127
#pragma unroll 1
      for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
        uint32_t r[32];
        __syncwarp();
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c, r);
        ptx::tmem_ld_wait();
        const int col0 = n_tile * BN + c;
        if (col0 >= g.N || row0 >= g.M) continue;   // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(stg + lane * kStgLd + j) =
              make_float4(__uint_as_float(r[j]) * e.alpha, __uint_as_float(r[j + 1]) * e.alpha,
                          __uint_as_float(r[j + 2]) * e.alpha, __uint_as_float(r[j + 3]) * e.alpha);
        __syncwarp();

        if (e.gm_S > 0) {
          // max over the gm_S rows of each patch on the fp32 accumulators (torch.max, utils.py:180,188), first
          // index wins; lane = column.  Biases are constant down a column, so they are added after the max.
          const int S = e.gm_S;
          const int col = col0 + lane;
          if (col < g.N) {
            const float badd = e.bias ? __ldg(e.bias + col) : 0.f;
            for (int gr = 0; gr < 32; gr += S) {
              if (row0 + gr >= g.M) break;
              float m = stg[gr * kStgLd + lane];
              int am = 0;
              for (int s2 = 1; s2 < S; ++s2) {
                const float x = stg[(gr + s2) * kStgLd + lane];
                if (x > m) { m = x; am = s2; }
              }
              m += badd;
              const size_t go = (size_t)((row0 + gr) / S) * e.gm_ld + col;
              if (e.gm_out_f32) e.gm_out_f32[go] = m;
              if (e.gm_out_bf16) reinterpret_cast<__nv_bfloat16 *>(e.gm_out_bf16)[go] = __float2bfloat16(m);
              if (e.gm_argmax) e.gm_argmax[go] = (uint8_t)am;
            }
          }
          if (!e.out) continue;
        }

        const int col = col0 + cq;
        if (col >= g.N) continue;
        const int nv = min(4, g.N - col);
        if (vec_ok && nv == 4 && (e.aux_mode == VPF_AUX_NONE || (e.ld_aux & 3) == 0)) {
          // ---------------- fast path: whole float4 quads.  All global loads of the chunk are issued up front
          // (one epilogue warp has nobody to hide its latency behind), then the math, then the stores.
          float4 t[8];
          bool ok[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            t[it] = *reinterpret_cast<const float4 *>(stg + (it * 4 + rsub) * kStgLd + cq);
            ok[it] = row0 + it * 4 + rsub < g.M;
          }
          float4 rv[8], rg[8];
          uint2 ax[8];
          if (e.mode == VPF_EPI_RESIDUAL) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) rv[it] = __ldg(reinterpret_cast<const float4 *>(e.resid + (size_t)(row0 + it * 4 + rsub) * e.ldc + col));
          }
          if (e.aux_mode != VPF_AUX_NONE) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) ax[it] = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const __nv_bfloat16 *>(e.aux) + (size_t)(row0 + it * 4 + rsub) * e.ld_aux + col));
          }
          if (e.rg_bias) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) rg[it] = __ldg(reinterpret_cast<const float4 *>(e.rg_bias + (size_t)((row0 + it * 4 + rsub) >> e.rg_shift) * e.rg_ld + col));
          }
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e.bias) b4 = __ldg(reinterpret_cast<const float4 *>(e.bias + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!ok[it]) continue;
            const long long grow = row0 + it * 4 + rsub;
            const size_t off = (size_t)grow * e.ldc + col;
            float v[4] = {t[it].x + b4.x, t[it].y + b4.y, t[it].z + b4.z, t[it].w + b4.w};
            if (e.rg_bias) { v[0] += rg[it].x; v[1] += rg[it].y; v[2] += rg[it].z; v[3] += rg[it].w; }
            if (e.out2) {
              uint2 pk;
              *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
              *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(v[2], v[3]);
              *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(e.out2) + off) = pk;
            }
            if (e.act == VPF_ACT_RELU) {
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
            } else if (e.act == VPF_ACT_GELU) {
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = gelu_f(v[q]);
            }
            if (e.aux_mode != VPF_AUX_NONE) {
              const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ax[it].x));
              const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ax[it].y));
              const float a[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
              for (int q = 0; q < 4; ++q)
                v[q] = e.aux_mode == VPF_AUX_GELU_GRAD ? v[q] * gelu_grad_f(a[q]) : (a[q] > 0.f ? v[q] : 0.f);
            }
            if (e.mode == VPF_EPI_STORE) {
              if (e.out_f32) {
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(e.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
              } else {
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
                *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(e.out) + off) = pk;
              }
            } else if (e.mode == VPF_EPI_RESIDUAL) {   // out_f32 = resid + dropout(v)   (Residual, partseg.py:208-213)
              if (drop_thr) {
                const uint32_t ebase = (uint32_t)((size_t)grow * g.N + col);
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = rng::keep(drop_key, ebase + q, drop_thr) ? v[q] * drop_scale : 0.f;
              }
              v[0] += rv[it].x; v[1] += rv[it].y; v[2] += rv[it].z; v[3] += rv[it].w;
              *reinterpret_cast<float4 *>(reinterpret_cast<float *>(e.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
              if (e.out_bf16) {
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
                *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(e.out_bf16) + off) = pk;
              }
            } else {   // VPF_EPI_ATOMIC_ADD: split-K weight gradients accumulate into the flat fp32 grad buffer
              ptx::red_add_v4(reinterpret_cast<float *>(e.out) + off, v[0], v[1], v[2], v[3]);
            }
          }
          continue;
        }
        // ---------------- generic path (ragged N, unaligned ld): scalar, element by element
        for (int it = 0; it < 8; ++it) {
          const int rl = it * 4 + rsub;
          const long long grow = row0 + rl;
          if (grow >= g.M) continue;
          for (int q = 0; q < nv; ++q) {
            float v = stg[rl * kStgLd + cq + q];
            const size_t off = (size_t)grow * e.ldc + col + q;
            if (e.bias) v += __ldg(e.bias + col + q);
            if (e.rg_bias) v += __ldg(e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + col + q);
            if (e.out2) reinterpret_cast<__nv_bfloat16 *>(e.out2)[off] = __float2bfloat16(v);
            if (e.act == VPF_ACT_RELU) v = fmaxf(v, 0.f);
            else if (e.act == VPF_ACT_GELU) v = gelu_f(v);
            if (e.aux_mode != VPF_AUX_NONE) {
              const float a = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(e.aux)[(size_t)grow * e.ld_aux + col + q]);
              v = e.aux_mode == VPF_AUX_GELU_GRAD ? v * gelu_grad_f(a) : (a > 0.f ? v : 0.f);
            }
            if (e.mode == VPF_EPI_STORE) {
              if (e.out_f32) reinterpret_cast<float *>(e.out)[off] = v;
              else reinterpret_cast<__nv_bfloat16 *>(e.out)[off] = __float2bfloat16(v);
            } else if (e.mode == VPF_EPI_RESIDUAL) {
              if (drop_thr) v = rng::keep(drop_key, (uint32_t)((size_t)grow * g.N + col + q), drop_thr) ? v * drop_scale : 0.f;
              v += e.resid[off];
              reinterpret_cast<float *>(e.out)[off] = v;
              if (e.out_bf16) reinterpret_cast<__nv_bfloat16 *>(e.out_bf16)[off] = __float2bfloat16(v);
            } else {
              atomicAdd(reinterpret_cast<float *>(e.out) + off, v);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 tensor map: array [outer][inner] with row stride ld elements, box {box_inner, box_outer}, 128B swizzle
int make_tmap_bf16(CUtensorMap *m, const void *base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15))
    return fail(VPF_EINVAL, "gemm operand must be 16-byte aligned with a 16-byte-multiple row stride (ld=%llu)", (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return VPF_OK;
}

}  // namespace vpf

using namespace vpf;

extern "C" int vpf_gemm_bf16(const void *A, int a_mn, int lda, const void *B, int b_mn, int ldb, int M, int N, int K,
                             int splits, const vpf_gemm_epilogue *epi, void *stream) {
  VPF_REQUIRE(A && B && epi && (epi->out || epi->gm_S > 0), "gemm: null pointer");
  VPF_REQUIRE(epi->gm_S == 0 || ((epi->gm_S & (epi->gm_S - 1)) == 0 && epi->gm_S <= 32 && M % epi->gm_S == 0 && (epi->gm_ld % 16) == 0 && epi->mode == VPF_EPI_STORE), "gemm: group-max epilogue needs S a power of two <= 32 dividing M, gm_ld %% 16 == 0");
  VPF_REQUIRE(M >= 0 && N >= 0 && K >= 1, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  VPF_REQUIRE(epi->mode == VPF_EPI_STORE || epi->mode == VPF_EPI_RESIDUAL || epi->mode == VPF_EPI_ATOMIC_ADD, "gemm: bad epilogue mode %d", epi->mode);
  VPF_REQUIRE(epi->mode != VPF_EPI_RESIDUAL || epi->resid, "gemm: residual epilogue needs resid");
  VPF_REQUIRE(splits == 1 || epi->mode == VPF_EPI_ATOMIC_ADD, "gemm: split-K needs the atomic-add epilogue");
  VPF_REQUIRE((size_t)M * (size_t)N < (1ull << 32) || epi->drop_p == 0.f, "gemm: dropout index space exceeds 2^32");
  if (M == 0 || N == 0) return VPF_OK;
  CUtensorMap ta, tb;
  if (!a_mn) VPF_TRY(make_tmap_bf16(&ta, A, K, M, lda, BK, BM));
  else VPF_TRY(make_tmap_bf16(&ta, A, M, K, lda, 64, BK));
  if (!b_mn) VPF_TRY(make_tmap_bf16(&tb, B, K, N, ldb, BK, BN));
  else VPF_TRY(make_tmap_bf16(&tb, B, N, K, ldb, 64, BK));
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.a_mn = a_mn ? 1 : 0; g.b_mn = b_mn ? 1 : 0;
  g.num_m_tiles = ceil_div(M, BM);
  g.num_n_tiles = ceil_div(N, BN);
  g.kblocks = ceil_div(K, BK);
  if (splits < 1) {  // auto: fill the machine
    const int tiles = g.num_m_tiles * g.num_n_tiles;
    splits = epi->mode == VPF_EPI_ATOMIC_ADD ? max(1, min(g.kblocks, (2 * num_sms()) / max(1, tiles))) : 1;
  }
  g.kblocks_per_split = ceil_div(g.kblocks, splits);
  g.splits = ceil_div(g.kblocks, g.kblocks_per_split);
  g.e = *epi;
  static bool attr_set = false;
  if (!attr_set) {
    VPF_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    attr_set = true;
  }
  const int total = g.num_m_tiles * g.num_n_tiles * g.splits;
  const int grid = min(total, num_sms());
  gemm_bf16_kernel<<<grid, kGemmThreads, kGemmSmem, (cudaStream_t)stream>>>(ta, tb, g);
  return check_launch("gemm_bf16_kernel");
}
