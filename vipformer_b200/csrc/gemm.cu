// gemm.cu -- the dense-contraction workhorse: bf16 x bf16 -> fp32 on tcgen05.
//
//   C[M,N] (op)= epilogue( alpha * sum_k A(m,k) * B(n,k) )
//
// Persistent, warp-specialised kernel, one CTA per SM (576 threads), template <BN, STAGES, EK>:
//   warp 0      TMA producer   cp.async.bulk.tensor 2D, SWIZZLE_128B, ring of (A,B) 64-wide k-blocks
//   warp 1      MMA issuer     tcgen05.mma cta_group::1, M=128 N=BN (128 or 256) K=16, accumulators in TMEM; two
//                              accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..17 epilogue       EK >= 0: warp-autonomous (epilogue_wa): every warp owns a 32-row x 32-column chunk end
//                              to end -- TMEM quadrant, private staging slots, its own bulk stores / reduce-adds and
//                              residual / aux bulk loads -- no group barrier in the steady state; the epilogue
//                              options of the common kinds are compile-time constants.
//                              EK < 0: lock-step fallback (unaligned outputs, out2, transposed pool): two groups of
//                              8 warps, each owning 64 of the tile's 128 columns, staged 128x64 boxes, one elected
//                              TMA issue per group.
//                              Either way a thread owns one accumulator ROW (tcgen05.ld 32x32b), results leave by
//                              cp.async.bulk.tensor store or cp.reduce.async.bulk.tensor .add (split-K weight
//                              gradients) so M/N tails are clipped by the TMA unit, and the per-patch max pool
//                              (utils.py:180,188) reads the staged fp32 chunk column-wise.
//
// Both operands may be K-major (row-major [rows][K]) or MN-major ([K][rows]); that covers the forward (X.W^T), the
// data gradient (dY.W) and the weight gradient (dY^T.X) of every Linear / 1x1-Conv on the ViPFormer hot path
// (vipformer/model/pointcloud/partseg.py:15-198, utils.py:144-189, classifier.py:25-50) without a single transpose.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "gelu.cuh"

namespace vpf {

constexpr int BM = 128, BK = 64;
constexpr int kTileBytesA = BM * BK * 2;
constexpr int kGemmThreads = 576;   // TMA warp + MMA warp + 16 epilogue warps
constexpr int kBoxBytes = 128 * 128;              // one staging box: 128 rows x 128 bytes (32 fp32 or 64 bf16 columns)
constexpr int kStgF32 = 4 * kBoxBytes;            // fp32 staging: 2 groups x 2 boxes of 32 columns     (64 KB)
constexpr int kStgAux = 2 * kBoxBytes;            // bf16 staging: 2 groups x 1 box of 64 columns       (32 KB)

// Tile width BN_T = 128 or 256, ST_T smem pipeline stages.  What is left of the 227 KB is the epilogue staging area,
// used as one or two buffers (GemmArgs::stg_*): with two, the bulk store of one tile drains while the next tile's
// accumulators are converted -- measured (tools/gemm_sweep.py) the single-buffered epilogue, not the mainloop, bounds
// every K <= 256 GEMM at ~1.9 us per 128x128 bf16 tile (0.9 us of conversion + 1.0 us of store drain, in series).
template <int BN_T, int ST_T>
struct Cfg {
  static constexpr int BN = BN_T;
  static constexpr int kStages = ST_T;
  static constexpr int kTileBytesB = BN_T * BK * 2;
  static constexpr int kStageBytes = kTileBytesA + kTileBytesB;
  static constexpr int kOffStg = kStages * kStageBytes;
  static constexpr int kStgBytes = BN_T == 256 ? 4 * kBoxBytes : (ST_T == 4 ? 6 * kBoxBytes : 8 * kBoxBytes);   // 64 / 96 / 128 KB
  static constexpr int kOffBar = kOffStg + kStgBytes;
  static constexpr int kSmem = kOffBar + 256 + 1024 /*align slack*/;
  static constexpr int kTmemCols = 2 * BN_T;
  static constexpr int kHalvesPerGroup = BN_T / 128;      // 64-column halves each epilogue group walks per tile
};
static_assert(Cfg<128, 4>::kSmem <= 232448 && Cfg<128, 3>::kSmem <= 232448 && Cfg<256, 3>::kSmem <= 232448, "smem budget");

struct GemmArgs {
  int M, N, K;
  int a_mn, b_mn;
  int num_m_tiles, num_n_tiles, kblocks, kblocks_per_split, splits;
  int tma_epi;   // 1: TMA epilogue (aligned outputs); 0: generic direct-global epilogue
  int m_fast;    // 1: consecutive work items walk the M tiles first (they share the B tile), else the N tiles first
  int nt_shift;  // log2(num_n_tiles) when that is a power of two (shift/mask instead of div/mod per tile), else -1
  // epilogue staging: stg_nbuf buffers of stg_buf_bytes; inside a buffer the fp32 boxes sit at 0, the others at these offsets
  int stg_nbuf, stg_buf_bytes, stg_off_bf, stg_off_out2, stg_off_aux;
  vpf_gemm_epilogue e;
};

// work item -> (split, m_tile, n_tile).  Every warp of the CTA evaluates this once per tile, so the common case
// (no split-K, power-of-two N tiles, N tiles fastest) must not cost three integer divisions.
__device__ __forceinline__ void tile_coords(const GemmArgs &g, int w, int &split, int &m_tile, int &n_tile) {
  int wt = w;
  split = 0;
  if (g.splits > 1) {
    const int tiles = g.num_m_tiles * g.num_n_tiles;
    split = w / tiles;
    wt = w - split * tiles;
  }
  if (g.m_fast) {
    n_tile = wt / g.num_m_tiles;
    m_tile = wt - n_tile * g.num_m_tiles;
  } else if (g.nt_shift >= 0) {
    m_tile = wt >> g.nt_shift;
    n_tile = wt & (g.num_n_tiles - 1);
  } else {
    m_tile = wt / g.num_n_tiles;
    n_tile = wt - m_tile * g.num_n_tiles;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *m, const void *smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 16-byte piece k (0..7) of row `row` inside a 128-byte-row, 128B-swizzled box
__device__ __forceinline__ uint32_t swz(int row, int k) { return (uint32_t)(row * 128 + ((k ^ (row & 7)) << 4)); }

#ifdef VPF_GEMM_TIMING
// experiment-only build (VPF_NVCC_EXTRA=-DVPF_GEMM_TIMING): cycles per epilogue phase of CTA 0 / warp 2 / lane 0
__device__ unsigned long long g_gemm_phase[16];
#define VPF_TCK(i) do { if (dbg) { const long long t_ = clock64(); dbg_acc[i] += (unsigned long long)(t_ - dbg_t); dbg_t = t_; } } while (0)
#else
#define VPF_TCK(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------------------------
// Warp-autonomous epilogue.  Measured with the phase timers below: the lock-step epilogue (two groups of 8 warps, two
// named barriers and one elected TMA issue per 128x64 half) costs ~1.9 us per 128x128 tile whatever K is -- every warp
// runs a ~340-instruction dependent chain at ~10 cycles per instruction while all its peers wait in the same phase, so
// for K <= 256 the tensor pipe idles.  Here each of the 16 epilogue warps owns a 32-row x 32-column chunk end to end:
// its own TMEM quadrant, its own staging slots in shared memory, its own bulk stores (and residual / aux bulk loads
// with a private mbarrier).  No CTA- or group-wide barrier remains in the steady state, so the warps drift apart and
// hide each other's latencies; the accumulator stage is released right after the TMEM load.
//   staging of warp w: stg_nbuf buffers of stg_buf_bytes at stg + w * stg_nbuf * stg_buf_bytes; inside a buffer the
//   fp32 chunk (32 rows x 128 B, 128B-swizzled) sits at 0, the bf16 chunks (32 rows x 64 B, linear) at stg_off_bf /
//   stg_off_aux.
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst_s, const CUtensorMap *m, uint32_t bar_s, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst_s), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_s), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap *m, uint32_t src_s, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_s), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_s(const CUtensorMap *m, uint32_t src_s, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_s), "r"(c0), "r"(c1) : "memory");
}

// warp-autonomous epilogue kinds: generic, bf16 store, fp32 store / split-K reduce, residual (+dropout), and the three
// Group2Emb convolutions (bf16 store + per-patch max, per-patch max only, bf16 store + per-group bias)
enum { EK_GENERIC = 0, EK_BF16 = 1, EK_F32 = 2, EK_RESID = 3, EK_POOL_OUT = 4, EK_POOL = 5, EK_RG = 6 };

template <int BN_T, int EK>
__device__ __forceinline__ void epilogue_wa(const GemmArgs &g, const CUtensorMap *tma_out, const CUtensorMap *tma_resid,
                                            const CUtensorMap *tma_aux, uint32_t stg_s, uint64_t *tmem_full,
                                            uint64_t *tmem_empty, uint64_t *wld_bar, uint32_t tmem_base, int warp, int lane,
                                            int total_work) {
  constexpr int BN = BN_T, kPasses = BN_T / 128;
  const vpf_gemm_epilogue &e = g.e;
  const int ew = warp - 2;
  const int quad = warp & 3;              // TMEM lane quadrant this warp may read
  const int cg = ew >> 2;                 // 32-column group inside a 128-column pass
  // EK_GENERIC reads every epilogue option at run time; the other kinds fix them at compile time, which removes the
  // option tests (and the dead code behind them) from the per-chunk instruction stream
  constexpr bool G = EK == EK_GENERIC;
  const bool pool = G ? e.gm_S > 0 : (EK == EK_POOL || EK == EK_POOL_OUT);
  const bool resid = G ? e.mode == VPF_EPI_RESIDUAL : EK == EK_RESID;
  const bool has_aux = G ? e.aux_mode != VPF_AUX_NONE : false;
  const bool need_ld = resid || has_aux;
  const bool out_is_f32 = G ? (e.mode != VPF_EPI_STORE || e.out_f32) : (EK == EK_F32 || EK == EK_RESID);
  const bool has_out = G ? e.out != nullptr : EK != EK_POOL;
  const bool has_alpha = G ? e.alpha != 1.0f : false;
  const bool has_rg = G ? e.rg_bias != nullptr : EK == EK_RG;
  const int act = G ? e.act : (int)VPF_ACT_NONE;
  const bool atomic = (G || EK == EK_F32) ? e.mode == VPF_EPI_ATOMIC_ADD : false;
  const uint32_t my_s = stg_s + (uint32_t)(ew * g.stg_nbuf * g.stg_buf_bytes);
  const uint32_t bar_s = ptx::smem_u32(&wld_bar[ew]);
  uint32_t drop_thr = 0, drop_key = 0;
  float drop_scale = 1.f;
  if (resid && e.drop_p > 0.f) {
    drop_thr = rng::threshold8(e.drop_p);
    drop_key = rng::make_key(e.seed_ptr ? *e.seed_ptr : 0ull, e.op_id);
    drop_scale = rng::scale8(drop_thr);
  }
  const bool bias_vec = (e.bias || has_rg) && (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) &&
                        (!has_rg || ((reinterpret_cast<uintptr_t>(e.rg_bias) & 15) == 0 && (e.rg_ld & 3) == 0));
  // this lane's row inside the chunk: fp32 piece k at f32_row + ((k ^ (lane & 7)) << 4), bf16 piece k at lane * 64 + k * 16
  const uint32_t f32_row = (uint32_t)lane * 128u, x7 = (uint32_t)(lane & 7);
  const uint32_t ld_bytes = (resid ? 4096u : 0u) + (has_aux ? 2048u : 0u);
  int step = 0, iter = 0;
  uint32_t ld_phase = 0;
#ifdef VPF_GEMM_TIMING
  const bool dbg = blockIdx.x == 0 && warp == 2 && lane == 0;
  unsigned long long dbg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long dbg_t = clock64();
#endif
  for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++iter) {
    int split, m_tile, n_tile;
    tile_coords(g, w, split, m_tile, n_tile);
    const int acc = iter & 1;
    const int row0 = m_tile * BM + quad * 32;            // first row of this warp's chunk
    const long long grow = (long long)row0 + lane;
#pragma unroll 1
    for (int ps = 0; ps < kPasses; ++ps) {
      const int ccol = (cg + 4 * ps) * 32;               // chunk column inside the tile (= TMEM column offset)
      const int col0 = n_tile * BN + ccol;
      const bool active = col0 < g.N;                    // warp-uniform
      const uint32_t buf = my_s + ((g.stg_nbuf == 2 && (step & 1)) ? (uint32_t)g.stg_buf_bytes : 0u);
      const bool full_chunk = col0 + 32 <= g.N;
      const bool vec_bias = bias_vec && full_chunk;
      VPF_TCK(0);
      if (active) {
        ++step;
        // the bulk stores that last read this buffer must have finished READING it
        if (lane == 0) { if (g.stg_nbuf == 2) bulk_wait_read1(); else bulk_wait_read0(); }
        __syncwarp();
        if (need_ld && lane == 0) {   // residual / aux chunk: fetched while the MMA of this tile may still be running
          ptx::mbar_arrive_expect_tx(&wld_bar[ew], ld_bytes);
          if (resid) tma_load_2d_s(buf, tma_resid, bar_s, col0, row0);
          if (has_aux) tma_load_2d_s(buf + (uint32_t)g.stg_off_aux, tma_aux, bar_s, col0, row0);
        }
      }
      VPF_TCK(1);
      if (ps == 0) {
        ptx::mbar_wait(&tmem_full[acc], (iter >> 1) & 1);
        ptx::tc_fence_after();
      }
      VPF_TCK(2);
      uint32_t r[32];
      if (active) {
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + ccol, r);
        ptx::tmem_ld_wait();
      }
      VPF_TCK(3);
      if (ps == kPasses - 1) {   // accumulators are in registers: the MMA warp may refill this stage
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      }
      VPF_TCK(4);
      if (!active) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (has_alpha) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
      }
      if (pool) {   // raw accumulators for the per-patch max (bias is added after the max)
#pragma unroll
        for (int k = 0; k < 8; ++k) sts128f(buf + f32_row + (((uint32_t)k ^ x7) << 4), v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
      if (has_out) {
        if (vec_bias) {
          if (e.bias) {
            const float4 *bp = reinterpret_cast<const float4 *>(e.bias + col0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 t = __ldg(bp + k);
              v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          }
          if (has_rg && grow < g.M) {
            const float4 *rp = reinterpret_cast<const float4 *>(e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + col0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 t = __ldg(rp + k);
              v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          }
        } else if (e.bias || has_rg) {
          const int ncols = min(32, g.N - col0);
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(e.bias + col0 + j);
          }
          if (has_rg && grow < g.M) {
            const float *rb = e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(rb + j);
          }
        }
        if (act == VPF_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (act == VPF_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
        }
        if (need_ld) {
          ptx::mbar_wait(&wld_bar[ew], ld_phase);
          ld_phase ^= 1;
        }
        if (has_aux) {
          const uint32_t arow = buf + (uint32_t)g.stg_off_aux + (uint32_t)lane * 64u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 pk = lds128(arow + k * 16);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __bfloat1622float2(h[q]);
              float &x0 = v[8 * k + 2 * q], &x1 = v[8 * k + 2 * q + 1];
              if (e.aux_mode == VPF_AUX_GELU_GRAD) { x0 *= gelu_grad_f(a.x); x1 *= gelu_grad_f(a.y); }
              else { x0 = a.x > 0.f ? x0 : 0.f; x1 = a.y > 0.f ? x1 : 0.f; }
            }
          }
        }
        if (resid) {   // out = resid + dropout(v)   (Residual, partseg.py:208-213), in place in the staging chunk
          if (drop_thr) {
            const uint32_t ebase = (uint32_t)((size_t)grow * g.N + col0);
            rng::drop_values<32>(v, drop_key, ebase, drop_thr, drop_scale, (g.N & 3) == 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t a = buf + f32_row + (((uint32_t)k ^ x7) << 4);
            const float4 rr = lds128f(a);
            sts128f(a, rr.x + v[4 * k], rr.y + v[4 * k + 1], rr.z + v[4 * k + 2], rr.w + v[4 * k + 3]);
          }
        } else if (out_is_f32) {
#pragma unroll
          for (int k = 0; k < 8; ++k) sts128f(buf + f32_row + (((uint32_t)k ^ x7) << 4), v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
          const uint32_t brow = buf + (uint32_t)g.stg_off_bf + (uint32_t)lane * 64u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * k + 2 * q], v[8 * k + 2 * q + 1]);
              pk[q] = *reinterpret_cast<const uint32_t *>(&h);
            }
            sts128(brow + k * 16, pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      VPF_TCK(5);
      ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the bulk-copy engine
      VPF_TCK(6);
      __syncwarp();
      if (lane == 0) {
        if (has_out) {
          if (atomic) tma_reduce_add_2d_s(tma_out, buf, col0, row0);
          else if (out_is_f32) tma_store_2d_s(tma_out, buf, col0, row0);
          else tma_store_2d_s(tma_out, buf + (uint32_t)g.stg_off_bf, col0, row0);
        }
        bulk_commit();
      }
      VPF_TCK(7);
      if (pool) {
        // per-patch max over gm_S rows on the fp32 accumulators, first index wins (torch.max, utils.py:180,188);
        // the patch rows all live in this warp's chunk: lane = column
        const int col = col0 + lane;
        if (col < g.N) {
          const float badd = e.bias ? __ldg(e.bias + col) : 0.f;
          const int S = e.gm_S;
          const uint32_t kk = (uint32_t)lane >> 2, sub = ((uint32_t)lane & 3u) * 4u;
          for (int r0 = 0; r0 < 32; r0 += S) {
            if (row0 + r0 >= g.M) break;
            float m = -INFINITY;
            int am = 0;
            if (S >= 8) {
              for (int rb = 0; rb < S; rb += 8) {
                const uint32_t base = buf + (uint32_t)(r0 + rb) * 128u + sub;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ptx::lds_f32(base + (uint32_t)i * 128u + ((kk ^ (uint32_t)i) << 4));
#pragma unroll
                for (int i = 0; i < 8; ++i) if (x[i] > m || (rb + i) == 0) { m = x[i]; am = rb + i; }
              }
            } else {
              for (int s2 = 0; s2 < S; ++s2) {
                const uint32_t rr = (uint32_t)(r0 + s2);
                const float x = ptx::lds_f32(buf + rr * 128u + ((kk ^ (rr & 7u)) << 4) + sub);
                if (x > m || s2 == 0) { m = x; am = s2; }
              }
            }
            m += badd;
            const size_t go = (size_t)((row0 + r0) / S) * e.gm_ld + col;
            if (e.gm_out_f32) e.gm_out_f32[go] = m;
            if (e.gm_out_bf16) reinterpret_cast<__nv_bfloat16 *>(e.gm_out_bf16)[go] = __float2bfloat16(m);
            if (e.gm_argmax) e.gm_argmax[go] = (uint8_t)am;
          }
        }
        __syncwarp();   // all lanes are done reading the chunk before the next step may overwrite it
      }
      VPF_TCK(8);
    }
  }
#ifdef VPF_GEMM_TIMING
  if (dbg) {
    for (int i = 0; i < 12; ++i) atomicAdd(&g_gemm_phase[i], dbg_acc[i]);
    atomicAdd(&g_gemm_phase[12], (unsigned long long)step);
  }
#endif
  if (lane == 0) bulk_wait0();   // all bulk stores complete before the CTA (and its smem) goes away
}


template <int BN_T, int ST_T, int EK>   // EK < 0: lock-step epilogue
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_out2,
                 const __grid_constant__ CUtensorMap tma_resid, const __grid_constant__ CUtensorMap tma_aux,
                 const GemmArgs g) {
  using C = Cfg<BN_T, ST_T>;
  constexpr int BN = C::BN, kStages = C::kStages, kTileBytesB = C::kTileBytesB, kStageBytes = C::kStageBytes;
  constexpr int kOffStg = C::kOffStg, kOffBar = C::kOffBar, kTmemCols = C::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kOffBar);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *tmem_full = empty_bar + kStages;
  uint64_t *tmem_empty = tmem_full + 2;
  uint64_t *ld_bar = tmem_empty + 2;   // [2]: residual/aux prefetch of each epilogue group (lock-step epilogue)
  uint64_t *wld_bar = ld_bar + 2;      // [16]: residual/aux prefetch of each epilogue warp (warp-autonomous epilogue)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wld_bar + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const vpf_gemm_epilogue &e = g.e;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    if (g.tma_epi) {
      if (e.out) ptx::prefetch_tmap(&tma_out);
      if (e.out2) ptx::prefetch_tmap(&tma_out2);
      if (e.mode == VPF_EPI_RESIDUAL) ptx::prefetch_tmap(&tma_resid);
      if (e.aux_mode != VPF_AUX_NONE) ptx::prefetch_tmap(&tma_aux);
    }
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 16); ptx::mbar_init(&ld_bar[s], 1); }
      for (int s = 0; s < 16; ++s) ptx::mbar_init(&wld_bar[s], 1);
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
      int split, m_tile, n_tile;
      tile_coords(g, w, split, m_tile, n_tile);
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
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * (64 * BK * 2), &tma_b, &full_bar[stage], n_tile * BN + j * 64, kb * BK);
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
    // MN-major tile: two 64-wide blocks 8 KB apart (LBO), 8-k-row groups 1024 B apart (SBO); one UMMA_K = +2048 B.
    const uint32_t a_lbo = g.a_mn ? 64 * BK * 2 : 16, b_lbo = g.b_mn ? 64 * BK * 2 : 16;
    const uint32_t a_adv = g.a_mn ? 2048 : 32, b_adv = g.b_mn ? 2048 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++iter) {
      const int split = g.splits > 1 ? w / (g.num_m_tiles * g.num_n_tiles) : 0;
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
  } else if constexpr (EK >= 0) {
    epilogue_wa<BN_T, EK>(g, &tma_out, &tma_resid, &tma_aux, ptx::smem_u32(smem + kOffStg), tmem_full, tmem_empty, wld_bar,
                      tmem_base, warp, lane, total_work);
  } else {
    // ---------------------------------------------------------------- lock-step epilogue (generic fallback)
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 3;        // column half (64 columns) of the tile this group of 8 warps owns
    const int c2 = ((warp - 2) >> 2) & 1;   // which 32-column chunk of that half this warp handles
    const int row = quad * 32 + lane;       // this thread's row inside the 128-row tile
    const bool leader = (warp - 2) == grp * 8 && lane == 0;
    const int bar_id = 1 + grp;
    constexpr int kGrpThreads = 256;
    const bool gm = e.gm_S > 0;
    const bool out_is_f32 = e.mode != VPF_EPI_STORE || e.out_f32;
    int hstep = 0;                           // active half-steps of this group so far (selects the staging buffer)
#ifdef VPF_GEMM_TIMING
    const bool dbg = blockIdx.x == 0 && warp == 2 && lane == 0;
    unsigned long long dbg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long dbg_t = clock64();
#endif
    const bool need_ld = (e.mode == VPF_EPI_RESIDUAL) || (e.aux_mode != VPF_AUX_NONE);
    uint32_t drop_thr = 0, drop_key = 0;
    float drop_scale = 1.f;
    if (e.mode == VPF_EPI_RESIDUAL && e.drop_p > 0.f) {
      drop_thr = rng::threshold8(e.drop_p);
      drop_key = rng::make_key(e.seed_ptr ? *e.seed_ptr : 0ull, e.op_id);
      drop_scale = rng::scale8(drop_thr);
    }
    // vector (16-byte) bias loads need aligned pointers; otherwise the scalar path after the wait is used
    const bool bias_vec = (e.bias || e.rg_bias) && (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) &&
                          (!e.rg_bias || ((reinterpret_cast<uintptr_t>(e.rg_bias) & 15) == 0 && (e.rg_ld & 3) == 0));
    int iter = 0;
    uint32_t ld_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++iter) {
      int split, m_tile, n_tile;
      tile_coords(g, w, split, m_tile, n_tile);
      const int acc = iter & 1;
      const int tile_row0 = m_tile * BM;
      const long long grow = (long long)tile_row0 + row;
      // the group walks its 64-column halves of the tile one after the other through the same staging boxes
      // (one half at BN 128, two at BN 256: halves grp and grp + 2)
#pragma unroll 1
      for (int hf = 0; hf < C::kHalvesPerGroup; ++hf) {
      const int half = grp + 2 * hf;
      const int colg0 = n_tile * BN + half * 64;           // first column of this group's current half
      const bool grp_active = colg0 < g.N;                 // uniform over the group
      // staging locations of this group's 64 columns inside the buffer of this half-step
      uint8_t *stg = smem + kOffStg + (g.stg_nbuf == 2 ? (hstep & 1) * g.stg_buf_bytes : 0);
      uint8_t *f32_box = stg + grp * 2 * kBoxBytes;                        // two fp32 boxes (2 x 32 columns)
      uint8_t *bf_out_box = stg + g.stg_off_bf + grp * kBoxBytes;          // one bf16 box (64 columns)
      uint8_t *bf_out2_box = stg + g.stg_off_out2 + grp * kBoxBytes;
      uint8_t *aux_box = stg + g.stg_off_aux + grp * kBoxBytes;
      const uint32_t f32_box_s = ptx::smem_u32(f32_box);
      if (grp_active) ++hstep;
      VPF_TCK(0);

      if (g.tma_epi && need_ld) {
        // residual / aux tiles are prefetched into the staging area while the MMA of this tile is still running;
        // the staging is free once the previous tile's bulk stores have READ it
        if (leader) { if (g.stg_nbuf == 2) bulk_wait_read1(); else bulk_wait_read0(); }
        named_bar_sync(bar_id, kGrpThreads);
        if (leader && grp_active) {
          uint32_t bytes = 0;
          if (e.mode == VPF_EPI_RESIDUAL) bytes += 2 * kBoxBytes;
          if (e.aux_mode != VPF_AUX_NONE) bytes += kBoxBytes;
          ptx::mbar_arrive_expect_tx(&ld_bar[grp], bytes);
          if (e.mode == VPF_EPI_RESIDUAL) {
            ptx::tma_load_2d(f32_box, &tma_resid, &ld_bar[grp], colg0, tile_row0);
            ptx::tma_load_2d(f32_box + kBoxBytes, &tma_resid, &ld_bar[grp], colg0 + 32, tile_row0);
          }
          if (e.aux_mode != VPF_AUX_NONE) ptx::tma_load_2d(aux_box, &tma_aux, &ld_bar[grp], colg0, tile_row0);
        }
      }
      // bias (+ row-group bias) of this thread's 32 columns, fetched while the MMA of this tile is still running
      const int pcol0 = colg0 + c2 * 32;
      const bool pre_bias = g.tma_epi && bias_vec && pcol0 + 32 <= g.N;
      float4 pb[8];
      if (pre_bias) {
        if (e.bias) {
          const float4 *bp = reinterpret_cast<const float4 *>(e.bias + pcol0);
#pragma unroll
          for (int k = 0; k < 8; ++k) pb[k] = __ldg(bp + k);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) pb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (e.rg_bias && grow < g.M) {
          const float4 *rp = reinterpret_cast<const float4 *>(e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + pcol0);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 t = __ldg(rp + k);
            pb[k].x += t.x; pb[k].y += t.y; pb[k].z += t.z; pb[k].w += t.w;
          }
        }
      }
      VPF_TCK(1);
      if (hf == 0) {
        ptx::mbar_wait(&tmem_full[acc], (iter >> 1) & 1);
        ptx::tc_fence_after();
      }
      VPF_TCK(2);
      if (g.tma_epi && need_ld && grp_active) {
        ptx::mbar_wait(&ld_bar[grp], ld_phase);
        ld_phase ^= 1;
      }
      if (g.tma_epi && !need_ld) {
        // nothing to prefetch: only now (after the accumulator wait, so the previous tile's bulk store has had the whole
        // main loop to drain) make sure the staging area has been read
        if (leader) { if (g.stg_nbuf == 2) bulk_wait_read1(); else bulk_wait_read0(); }
        named_bar_sync(bar_id, kGrpThreads);
      }

      VPF_TCK(3);
      do {
        uint32_t r[32];
        __syncwarp();
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + half * 64 + c2 * 32, r);
        ptx::tmem_ld_wait();
        VPF_TCK(4);
        const int col0 = colg0 + c2 * 32;
        if (col0 >= g.N && !g.tma_epi) continue;   // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (e.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
        }

        if (!g.tma_epi) {
          // ------------- generic epilogue (unaligned / odd strides): direct global access, element by element
          if (grow < g.M) {
            const int ncols = min(32, g.N - col0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j >= ncols) continue;
              float x = v[j];
              const size_t off = (size_t)grow * e.ldc + col0 + j;
              if (e.bias) x += __ldg(e.bias + col0 + j);
              if (e.rg_bias) x += __ldg(e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + col0 + j);
              if (e.out2) reinterpret_cast<__nv_bfloat16 *>(e.out2)[off] = __float2bfloat16(x);
              if (e.act == VPF_ACT_RELU) x = fmaxf(x, 0.f);
              else if (e.act == VPF_ACT_GELU) x = gelu_f(x);
              if (e.aux_mode != VPF_AUX_NONE) {
                const float a = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(e.aux)[(size_t)grow * e.ld_aux + col0 + j]);
                x = e.aux_mode == VPF_AUX_GELU_GRAD ? x * gelu_grad_f(a) : (a > 0.f ? x : 0.f);
              }
              if (e.mode == VPF_EPI_STORE) {
                if (e.out_f32) reinterpret_cast<float *>(e.out)[off] = x;
                else reinterpret_cast<__nv_bfloat16 *>(e.out)[off] = __float2bfloat16(x);
              } else if (e.mode == VPF_EPI_RESIDUAL) {
                if (drop_thr) x = rng::keep8(drop_key, (uint32_t)((size_t)grow * g.N + col0 + j), drop_thr) ? x * drop_scale : 0.f;
                reinterpret_cast<float *>(e.out)[off] = x + e.resid[off];
              } else {
                atomicAdd(reinterpret_cast<float *>(e.out) + off, x);
              }
            }
          }
          continue;
        }
        if (!grp_active) continue;

        if (gm && e.gm_cols) {
          // transposed call: this thread owns ONE CHANNEL (row of C) and 32 consecutive points (columns) = whole
          // patches: the max pool (utils.py:180,188) is pure register work, first index wins, no staging at all
          if (grow < g.M) {
            const float badd = e.row_bias ? __ldg(e.row_bias + grow) : 0.f;
            const int S = e.gm_S, smask = S - 1;   // S: power of two <= 32
            float mx = v[0];
            int am = 0;
#pragma unroll
            for (int j = 1; j <= 32; ++j) {
              if (j == 32 || (j & smask) == 0) {          // patch boundary: emit the finished patch
                const int gs = j - S;
                if (col0 + gs < g.N) {
                  const size_t go = (size_t)((col0 + gs) / S) * e.gm_ld + grow;
                  const float o = mx + badd;
                  if (e.gm_out_f32) e.gm_out_f32[go] = o;
                  if (e.gm_out_bf16) reinterpret_cast<__nv_bfloat16 *>(e.gm_out_bf16)[go] = __float2bfloat16(o);
                  if (e.gm_argmax) e.gm_argmax[go] = (uint8_t)am;
                }
                if (j < 32) { mx = v[j]; am = 0; }
              } else if (v[j] > mx) {
                mx = v[j];
                am = j & smask;
              }
            }
          }
          continue;
        }
        // ------------- TMA epilogue: registers -> swizzled staging (this thread = one row).  Columns >= N inside
        // the box hold garbage/zeros; the TMA store clips them.
        if (gm) {   // raw accumulators for the column-wise max pool (biases are added after the max)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4 *>(f32_box + c2 * kBoxBytes + swz(row, k)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (!e.out) continue;
        }
        const int ncols = min(32, g.N - col0);   // may be <= 0 for the second chunk of a ragged tile
        if (pre_bias) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { v[4 * k] += pb[k].x; v[4 * k + 1] += pb[k].y; v[4 * k + 2] += pb[k].z; v[4 * k + 3] += pb[k].w; }
        } else {
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(e.bias + col0 + j);
          }
          if (e.rg_bias && grow < g.M) {
            const float *rb = e.rg_bias + (size_t)(grow >> e.rg_shift) * e.rg_ld + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(rb + j);
          }
        }
        if (e.out2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 pk;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[8 * k + 2 * q], v[8 * k + 2 * q + 1]);
            *reinterpret_cast<uint4 *>(bf_out2_box + swz(row, c2 * 4 + k)) = pk;
          }
        }
        if (e.act == VPF_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (e.act == VPF_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
        }
        if (e.aux_mode != VPF_AUX_NONE) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 pk = *reinterpret_cast<const uint4 *>(aux_box + swz(row, c2 * 4 + k));
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __bfloat1622float2(h[q]);
              float &x0 = v[8 * k + 2 * q], &x1 = v[8 * k + 2 * q + 1];
              if (e.aux_mode == VPF_AUX_GELU_GRAD) { x0 *= gelu_grad_f(a.x); x1 *= gelu_grad_f(a.y); }
              else { x0 = a.x > 0.f ? x0 : 0.f; x1 = a.y > 0.f ? x1 : 0.f; }
            }
          }
        }
        if (e.mode == VPF_EPI_RESIDUAL) {   // out = resid + dropout(v)   (Residual, partseg.py:208-213); in place in smem
          if (drop_thr) {
            const uint32_t ebase = (uint32_t)((size_t)grow * g.N + col0);
            rng::drop_values<32>(v, drop_key, ebase, drop_thr, drop_scale, (g.N & 3) == 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 *p = reinterpret_cast<float4 *>(f32_box + c2 * kBoxBytes + swz(row, k));
            const float4 rr = *p;
            *p = make_float4(rr.x + v[4 * k], rr.y + v[4 * k + 1], rr.z + v[4 * k + 2], rr.w + v[4 * k + 3]);
          }
        } else if (out_is_f32) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4 *>(f32_box + c2 * kBoxBytes + swz(row, k)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 pk;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[8 * k + 2 * q], v[8 * k + 2 * q + 1]);
            *reinterpret_cast<uint4 *>(bf_out_box + swz(row, c2 * 4 + k)) = pk;
          }
        }
      } while (0);
      VPF_TCK(5);
      // accumulator stage can be refilled by the MMA warp once the last half has been read
      if (hf == C::kHalvesPerGroup - 1) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      }

      VPF_TCK(6);
      if (g.tma_epi && grp_active) {
        ptx::fence_proxy_async();            // make the generic-proxy smem writes visible to the TMA unit
        VPF_TCK(7);
        named_bar_sync(bar_id, kGrpThreads);
        VPF_TCK(8);
        if (leader) {
          if (e.out) {
            if (e.mode == VPF_EPI_ATOMIC_ADD) {
              tma_reduce_add_2d(&tma_out, f32_box, colg0, tile_row0);
              if (colg0 + 32 < g.N) tma_reduce_add_2d(&tma_out, f32_box + kBoxBytes, colg0 + 32, tile_row0);
            } else if (out_is_f32) {
              tma_store_2d(&tma_out, f32_box, colg0, tile_row0);
              if (colg0 + 32 < g.N) tma_store_2d(&tma_out, f32_box + kBoxBytes, colg0 + 32, tile_row0);
            } else {
              tma_store_2d(&tma_out, bf_out_box, colg0, tile_row0);
            }
          }
          if (e.out2) tma_store_2d(&tma_out2, bf_out2_box, colg0, tile_row0);
          bulk_commit();
        }
        VPF_TCK(9);
        if (gm && !e.gm_cols) {
          // per-patch max over gm_S rows on the fp32 accumulators, first index wins (torch.max, utils.py:180,188).
          // thread t of the group: column t % 64, row quarter t / 64.
          const int t = (warp - 2 - grp * 8) * 32 + lane;
          const int cl = t & 63, col = colg0 + cl;
          if (col < g.N) {
            const float badd = e.bias ? __ldg(e.bias + col) : 0.f;
            const uint8_t *box = f32_box + (cl >> 5) * kBoxBytes;
            const int k = (cl & 31) >> 2, sub = (cl & 3) * 4;
            const int S = e.gm_S;
            const uint32_t col_s = f32_box_s + (cl >> 5) * kBoxBytes + sub;   // shared-window address of (row 0, this column)
            uint32_t off[8];                                                  // + row i of an aligned 8-row block
#pragma unroll
            for (int i = 0; i < 8; ++i) off[i] = (uint32_t)(i * 128 + ((k ^ i) << 4));
            for (int r0 = (t >> 6) * 32; r0 < (t >> 6) * 32 + 32; r0 += S) {
              if (tile_row0 + r0 >= g.M) break;
              float m;
              int am = 0;
              if (S >= 8) {
                m = -INFINITY;
                for (int rb = 0; rb < S; rb += 8) {
                  const uint32_t base = col_s + (uint32_t)(r0 + rb) * 128;
                  float x[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) x[i] = ptx::lds_f32(base + off[i]);
#pragma unroll
                  for (int i = 0; i < 8; ++i) if (x[i] > m || (rb + i) == 0) { m = x[i]; am = rb + i; }
                }
              } else {
                m = *reinterpret_cast<const float *>(box + swz(r0, k) + sub);
                for (int s2 = 1; s2 < S; ++s2) {
                  const float x = *reinterpret_cast<const float *>(box + swz(r0 + s2, k) + sub);
                  if (x > m) { m = x; am = s2; }
                }
              }
              m += badd;
              const size_t go = (size_t)((tile_row0 + r0) / S) * e.gm_ld + col;
              if (e.gm_out_f32) e.gm_out_f32[go] = m;
              if (e.gm_out_bf16) reinterpret_cast<__nv_bfloat16 *>(e.gm_out_bf16)[go] = __float2bfloat16(m);
              if (e.gm_argmax) e.gm_argmax[go] = (uint8_t)am;
            }
          }
        }
      }
      VPF_TCK(10);
      }   // halves
    }
#ifdef VPF_GEMM_TIMING
    if (dbg) {
      for (int i = 0; i < 12; ++i) atomicAdd(&g_gemm_phase[i], dbg_acc[i]);
      atomicAdd(&g_gemm_phase[12], (unsigned long long)hstep);
    }
#endif
    if (g.tma_epi && leader) bulk_wait0();   // all bulk stores complete before the CTA (and its smem) goes away
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

// 2D tensor map: array [outer][inner] with row stride ld elements, box {box_inner, box_outer}, 128B swizzle
static int make_tmap(CUtensorMap *m, const void *base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer, bool swizzle128 = true) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15))
    return fail(VPF_EINVAL, "gemm operand must be 16-byte aligned with a 16-byte-multiple row stride (ld=%llu)", (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VPF_ECUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return VPF_OK;
}

static bool tma_ok(const void *p, long long ld, int elem_bytes) {
  return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * elem_bytes) & 15) == 0);
}

}  // namespace vpf

using namespace vpf;

extern "C" int vpf_gemm_bf16(const void *A, int a_mn, int lda, const void *B, int b_mn, int ldb, int M, int N, int K,
                             int splits, const vpf_gemm_epilogue *epi, void *stream) {
  VPF_REQUIRE(A && B && epi && (epi->out || epi->gm_S > 0), "gemm: null pointer");
  VPF_REQUIRE(epi->gm_S == 0 || ((epi->gm_S & (epi->gm_S - 1)) == 0 && epi->gm_S <= 32 && (epi->gm_cols ? N : M) % epi->gm_S == 0 && epi->mode == VPF_EPI_STORE && epi->alpha == 1.0f),
              "gemm: max-pool epilogue needs S a power of two <= 32 dividing the pooled dimension, store mode, alpha 1");
  VPF_REQUIRE(!epi->gm_cols || (epi->gm_S > 0 && epi->out == nullptr), "gemm: gm_cols is a pool-only epilogue (out must be NULL)");
  VPF_REQUIRE(M >= 0 && N >= 0 && K >= 1, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  VPF_REQUIRE(epi->mode == VPF_EPI_STORE || epi->mode == VPF_EPI_RESIDUAL || epi->mode == VPF_EPI_ATOMIC_ADD, "gemm: bad epilogue mode %d", epi->mode);
  VPF_REQUIRE(epi->mode != VPF_EPI_RESIDUAL || epi->resid, "gemm: residual epilogue needs resid");
  VPF_REQUIRE(epi->out_bf16 == nullptr, "gemm: out_bf16 is no longer supported");
  VPF_REQUIRE(epi->out2 == nullptr || (epi->mode == VPF_EPI_STORE && !epi->out_f32 && epi->gm_S == 0), "gemm: out2 needs a bf16 store epilogue without max pool");
  VPF_REQUIRE(epi->gm_S == 0 || epi->aux_mode == VPF_AUX_NONE, "gemm: max-pool epilogue cannot be combined with aux");
  VPF_REQUIRE(splits <= 1 || epi->mode == VPF_EPI_ATOMIC_ADD, "gemm: split-K needs the atomic-add epilogue");
  VPF_REQUIRE((size_t)M * (size_t)N < (1ull << 32) || epi->drop_p == 0.f, "gemm: dropout index space exceeds 2^32");
  if (M == 0 || N == 0) return VPF_OK;
  GemmArgs g;
  g.e = *epi;
  const bool out_f32 = epi->mode != VPF_EPI_STORE || epi->out_f32;
  g.tma_epi = tma_ok(epi->out, epi->ldc, out_f32 ? 4 : 2) && tma_ok(epi->out2, epi->ldc, 2) &&
              tma_ok(epi->mode == VPF_EPI_RESIDUAL ? epi->resid : nullptr, epi->ldc, 4) &&
              tma_ok(epi->aux_mode != VPF_AUX_NONE ? epi->aux : nullptr, epi->ld_aux, 2);
  VPF_REQUIRE(g.tma_epi || epi->gm_S == 0, "gemm: max-pool epilogue needs 16-byte aligned outputs");
  // ---- epilogue flavour, staging layout and kernel variant
  const bool pool = epi->gm_S > 0 && !epi->gm_cols;
  const bool has_aux = epi->aux_mode != VPF_AUX_NONE;
  const bool wa = g.tma_epi && !epi->out2 && !epi->gm_cols;   // warp-autonomous epilogue (everything the step uses)
  // measured per shape class (tools/gemm_table.py, tools/gemm_sweep.py) with the warp-autonomous epilogue: the wide
  // tile (25 % less shared-memory traffic per FLOP) wins for K >= 256 and for bf16 outputs at any K; fp32 outputs with
  // K < 256 stay narrow with two staging buffers; split-K wants pipeline depth.
  // VPF_GEMM_BN=128|256 overrides the tile width (experiments).
  static const int force_bn = [] { const char *v = getenv("VPF_GEMM_BN"); return v ? atoi(v) : 0; }();
  int variant = 0;   // 0: <128,4>  1: <128,3>  2: <256,3>
  g.stg_off_bf = g.stg_off_out2 = g.stg_off_aux = 0;
  if (wa) {
    int cb = 0;   // bytes of one staging buffer of one warp (32 x 32 chunk)
    if (out_f32 || pool) cb = 4096;                                    // fp32 chunk (output, residual or pool input)
    if (epi->out && !out_f32) { g.stg_off_bf = cb; cb += 2048; }       // bf16 output chunk
    if (has_aux) { g.stg_off_aux = cb; cb += 2048; }                   // bf16 aux chunk
    g.stg_buf_bytes = cb;
    const bool wide_ok = N % 256 == 0 && 16 * cb <= Cfg<256, 3>::kStgBytes;
    // (split-K: the wide tile halves the tile count, which only pays when enough tiles are left to spread over the SMs)
    bool wide = wide_ok && (epi->mode == VPF_EPI_ATOMIC_ADD ? (long long)M * N >= 512LL * 256 : (K >= 256 || !out_f32));
    if (force_bn == 128) wide = false;
    if (force_bn == 256) wide = wide_ok;
    if (wide) variant = 2;
    else if (32 * cb <= Cfg<128, 4>::kStgBytes || epi->mode == VPF_EPI_ATOMIC_ADD || K >= 512) variant = 0;
    else variant = 1;
    const int cap = variant == 2 ? Cfg<256, 3>::kStgBytes : (variant == 0 ? Cfg<128, 4>::kStgBytes : Cfg<128, 3>::kStgBytes);
    VPF_REQUIRE(16 * cb <= cap, "gemm: epilogue staging of %d bytes per warp does not fit", cb);
    g.stg_nbuf = 32 * cb <= cap ? 2 : 1;
  } else {
    int sb = 0;   // lock-step epilogue: one buffer for the whole tile, fp32 boxes at 0
    if (out_f32 || pool) sb = 4 * kBoxBytes;
    if (epi->out && !out_f32) { g.stg_off_bf = sb; sb += 2 * kBoxBytes; }
    if (epi->out2) { g.stg_off_out2 = sb; sb += 2 * kBoxBytes; }
    if (has_aux) { g.stg_off_aux = sb; sb += 2 * kBoxBytes; }
    if (sb == 0) sb = 2 * kBoxBytes;
    constexpr int cap0 = Cfg<128, 4>::kStgBytes;
    VPF_REQUIRE(sb <= cap0, "gemm: epilogue staging of %d bytes does not fit", sb);
    g.stg_buf_bytes = sb;
    g.stg_nbuf = 1;
  }
  const int BN = variant == 2 ? 256 : 128;
  CUtensorMap ta, tb, tout, tout2, tres, taux;
  if (!a_mn) VPF_TRY(make_tmap(&ta, A, 2, K, M, lda, BK, BM));
  else VPF_TRY(make_tmap(&ta, A, 2, M, K, lda, 64, BK));
  if (!b_mn) VPF_TRY(make_tmap(&tb, B, 2, K, N, ldb, BK, BN));
  else VPF_TRY(make_tmap(&tb, B, 2, N, K, ldb, 64, BK));
  tout = tout2 = tres = taux = ta;   // placeholders for unused maps
  if (wa) {   // 32 x 32 chunks: fp32 rows are 128 B (swizzled), bf16 rows 64 B (linear)
    if (epi->out) VPF_TRY(make_tmap(&tout, epi->out, out_f32 ? 4 : 2, N, M, epi->ldc, 32, 32, out_f32));
    if (epi->mode == VPF_EPI_RESIDUAL) VPF_TRY(make_tmap(&tres, epi->resid, 4, N, M, epi->ldc, 32, 32));
    if (has_aux) VPF_TRY(make_tmap(&taux, epi->aux, 2, N, M, epi->ld_aux, 32, 32, false));
  } else if (g.tma_epi) {
    if (epi->out) VPF_TRY(make_tmap(&tout, epi->out, out_f32 ? 4 : 2, N, M, epi->ldc, out_f32 ? 32 : 64, BM));
    if (epi->out2) VPF_TRY(make_tmap(&tout2, epi->out2, 2, N, M, epi->ldc, 64, BM));
    if (epi->mode == VPF_EPI_RESIDUAL) VPF_TRY(make_tmap(&tres, epi->resid, 4, N, M, epi->ldc, 32, BM));
    if (has_aux) VPF_TRY(make_tmap(&taux, epi->aux, 2, N, M, epi->ld_aux, 64, BM));
  }
  g.M = M; g.N = N; g.K = K; g.a_mn = a_mn ? 1 : 0; g.b_mn = b_mn ? 1 : 0;
  g.num_m_tiles = ceil_div(M, BM);
  g.num_n_tiles = ceil_div(N, BN);
  g.kblocks = ceil_div(K, BK);
  g.nt_shift = -1;
  for (int sft = 0; sft < 16; ++sft) if ((1 << sft) == g.num_n_tiles) g.nt_shift = sft;
  g.m_fast = (epi->gm_cols && N > M) ? 1 : 0;   // transposed pool: adjacent work items share the (large) activation tile
  if (splits < 1) {  // auto: fill the machine
    const int tiles = g.num_m_tiles * g.num_n_tiles;
    splits = epi->mode == VPF_EPI_ATOMIC_ADD ? max(1, min(g.kblocks, (2 * num_sms()) / max(1, tiles))) : 1;
  }
  g.kblocks_per_split = ceil_div(g.kblocks, splits);
  g.splits = ceil_div(g.kblocks, g.kblocks_per_split);
  // epilogue kind: the three shapes that make up the training step are compiled without run-time option tests
  int ek = EK_GENERIC;
  if (wa && epi->alpha == 1.0f && epi->act == VPF_ACT_NONE && !has_aux) {
    if (pool) {
      if (!epi->rg_bias && (!epi->out || !out_f32)) ek = epi->out ? EK_POOL_OUT : EK_POOL;
    } else if (epi->rg_bias) {
      if (epi->mode == VPF_EPI_STORE && !out_f32) ek = EK_RG;
    } else if (epi->mode == VPF_EPI_RESIDUAL) ek = EK_RESID;
    else if (out_f32) ek = EK_F32;
    else ek = EK_BF16;
  }
  const int total = g.num_m_tiles * g.num_n_tiles * g.splits;
  const int grid = min(total, num_sms());
  cudaStream_t st = (cudaStream_t)stream;
#define VPF_LAUNCH(BNV, STV, EKV)                                                                                     \
  do {                                                                                                                \
    static bool attr_done = false;                                                                                    \
    if (!attr_done) {                                                                                                 \
      VPF_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_kernel<BNV, STV, EKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        Cfg<BNV, STV>::kSmem));                                                       \
      attr_done = true;                                                                                               \
    }                                                                                                                 \
    gemm_bf16_kernel<BNV, STV, EKV><<<grid, kGemmThreads, Cfg<BNV, STV>::kSmem, st>>>(ta, tb, tout, tout2, tres, taux, g); \
  } while (0)
#define VPF_LAUNCH_EK(BNV, STV)                                     \
  do {                                                              \
    if (ek == EK_BF16) VPF_LAUNCH(BNV, STV, EK_BF16);               \
    else if (ek == EK_F32) VPF_LAUNCH(BNV, STV, EK_F32);            \
    else if (ek == EK_RESID) VPF_LAUNCH(BNV, STV, EK_RESID);        \
    else if (ek == EK_POOL_OUT) VPF_LAUNCH(BNV, STV, EK_POOL_OUT);  \
    else if (ek == EK_POOL) VPF_LAUNCH(BNV, STV, EK_POOL);          \
    else if (ek == EK_RG) VPF_LAUNCH(BNV, STV, EK_RG);              \
    else VPF_LAUNCH(BNV, STV, EK_GENERIC);                          \
  } while (0)
  if (!wa) VPF_LAUNCH(128, 4, -1);
  else if (variant == 2) VPF_LAUNCH_EK(256, 3);
  else if (variant == 1) VPF_LAUNCH_EK(128, 3);
  else VPF_LAUNCH_EK(128, 4);
#undef VPF_LAUNCH_EK
#undef VPF_LAUNCH
  return check_launch("gemm_bf16_kernel");
}

#ifdef VPF_GEMM_TIMING
// experiment-only: copy out and clear the phase counters
extern "C" int vpf_debug_gemm_phase(unsigned long long *out16) {
  VPF_CUDA_TRY(cudaDeviceSynchronize());
  VPF_CUDA_TRY(cudaMemcpyFromSymbol(out16, g_gemm_phase, sizeof(unsigned long long) * 16));
  unsigned long long z[16] = {0};
  VPF_CUDA_TRY(cudaMemcpyToSymbol(g_gemm_phase, z, sizeof(z)));
  return VPF_OK;
}
#endif
