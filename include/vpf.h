/*
 * vpf.h -- C ABI of libvpf_b200.so, the B200 (sm_100a) implementation of the
 * ViPFormer pre-training hot path.
 *
 * The reference has no FFI: its boundary is a set of Python callables
 * (SURVEY.md section 8b).  Each entry below names the reference callable it
 * sits under (paths relative to the upstream repo root); the Python mirror in
 * vipformer_b200/ binds these with ctypes and keeps the reference signatures.
 *
 * Conventions (every entry):
 *   - plain pointers + sizes, no torch / C++ types;
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - row-major, contiguous; fp32 = float, bf16 = uint16_t storage,
 *     indices int64_t (what the reference returns);
 *   - `stream` is a cudaStream_t passed as void*; entries enqueue work on it
 *     and return without synchronising; they never allocate device memory
 *     (scratch is caller-provided through the *_workspace_bytes queries);
 *   - return 0 on success, a negative VPF_E* code otherwise;
 *     vpf_last_error_string() describes the last failure on this host thread;
 *   - there is NO CPU fallback anywhere in this library.
 */
#ifndef VPF_H_
#define VPF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPF_OK 0
#define VPF_EINVAL (-1)      /* bad argument / unsupported shape */
#define VPF_ECUDA (-2)       /* a CUDA runtime / driver call failed */
#define VPF_EWORKSPACE (-3)  /* workspace too small */

#define VPF_ABI_VERSION 1

int vpf_abi_version(void);
const char *vpf_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's
 * `gpu_launches` claim is read from here). */
int64_t vpf_launch_count(void);
void vpf_launch_count_reset(void);

/* ------------------------------------------------------------------ tokenizer
 * vipformer/model/pointcloud/utils.py */

/* farthest_point_sample, utils.py:56-85.  pts [B,N,C] fp32 (C>=3, channels
 * 0..2 used), start_idx [B] int64 (replaces the torch.randint draw of
 * utils.py:71), out_idx [B,npoint] int64.  Tie-break: lowest index.
 * Limits: 1 <= N <= 8192. */
int vpf_fps(const float *pts, int B, int N, int C, int npoint,
            const int64_t *start_idx, int64_t *out_idx, void *stream);

/* index_points, utils.py:88-104.  out[b,s,:] = points[b, idx[b,s], :]. */
int vpf_index_points(const float *points, int B, int N, int C,
                     const int64_t *idx, int S, float *out, void *stream);

/* square_distance, utils.py:122-141.  src [B,S,Cs], dst [B,N,Cd] (first three
 * channels), out [B,S,N] fp32, arithmetic pinned as in oracle/tokenizer_oracle.c */
int vpf_square_distance(const float *src, int B, int S, int Cs,
                        const float *dst, int N, int Cd, float *out,
                        void *stream);

/* knn_point, utils.py:107-119.  xyz [B,N,C], new_xyz [B,S,Cq] -> out_idx
 * [B,S,nsample] int64: the nsample smallest by (distance asc, index asc), in
 * that order.  Limits: nsample <= 32, nsample <= N <= 8192. */
int vpf_knn_point(int nsample, const float *xyz, int B, int N, int C,
                  const float *new_xyz, int S, int Cq, int64_t *out_idx,
                  void *stream);

/* divide_patches, utils.py:6-38 (FPS -> kNN -> gather -> slot-0..2 centre
 * subtraction, utils.py:36).  neighbors [B,G,S,C], centers [B,G,C];
 * fps_idx [B,G] / knn_idx [B,G,S] int64 are optional outputs (NULL to skip).
 * Limits: S <= 32, S <= N <= 8192, C >= 3. */
int vpf_divide_patches(const float *pts, int B, int N, int C, int G, int S,
                       const int64_t *start_idx, float *neighbors,
                       float *centers, int64_t *fps_idx, int64_t *knn_idx,
                       void *stream);

/* Same, HOST buffers in and out (pinned or pageable): H2D copy, the two
 * kernels, D2H copy, stream-synchronised before returning.  `workspace` is a
 * DEVICE scratch of at least vpf_divide_patches_host_workspace_bytes(). */
size_t vpf_divide_patches_host_workspace_bytes(int B, int N, int C, int G, int S);
int vpf_divide_patches_host(const float *pts_host, int B, int N, int C, int G,
                            int S, const int64_t *start_idx_host,
                            float *neighbors_host, float *centers_host,
                            void *workspace, size_t workspace_bytes,
                            void *stream);


/* Device-side augmentation: the trans_1 / trans_2 chain of datasets/data.py:16-36 (data_utils.py:56-221) on a batch of
 * clouds, in front of the tokenizer.  pts / out [B,N,3] fp32; params [B,6] = (scale, rotation angle about y, three
 * translation fractions of the bounding box, drop ratio); jitter [B,N,3] = N(0, std) draws before the clamp to
 * +-jitter_clip; drop_u [B,N] uniform draws (point i is overwritten by point 0 when drop_u <= drop ratio). */
int vpf_augment_clouds(const float *pts, const float *params, const float *jitter, const float *drop_u, float *out,
                       int B, int N, float jitter_clip, void *stream);

/* ------------------------------------------------------------ dense contraction
 * One tcgen05/TMA GEMM serves every Linear / 1x1-Conv1d on the path, forward and
 * backward (vipformer/model/pointcloud/partseg.py:46-49,191-198 Linear layers;
 * utils.py:153-165 Conv1d(k=1); classifier.py:31-36).
 *
 *   C[M,N] (op)= epilogue( alpha * sum_k A(m,k) * B(n,k) ),  bf16 operands, fp32 accumulate
 *
 * a_mn = 0: A is row-major [M][K] (row stride lda elements);  a_mn = 1: A is [K][M].
 * b_mn = 0: B is row-major [N][K] (row stride ldb elements);  b_mn = 1: B is [K][N].
 * splits: split-K factor (only with VPF_EPI_ATOMIC_ADD); 0 = choose automatically.
 * Operand base pointers must be 16-byte aligned, lda/ldb multiples of 8. */
#define VPF_EPI_STORE 0       /* out = f(acc)                         (bf16 or fp32)       */
#define VPF_EPI_RESIDUAL 1    /* out_f32 = resid + dropout(f(acc))    (Residual, partseg.py:201-213) */
#define VPF_EPI_ATOMIC_ADD 2  /* out_f32 += f(acc)  (red.global.add)  weight gradients     */
#define VPF_ACT_NONE 0
#define VPF_ACT_RELU 1
#define VPF_ACT_GELU 2        /* exact erf GELU (nn.GELU default) */
#define VPF_AUX_NONE 0
#define VPF_AUX_GELU_GRAD 1   /* f *= gelu'(aux)    (aux = saved pre-activation, bf16)      */
#define VPF_AUX_RELU_MASK 2   /* f  = aux > 0 ? f : 0                                       */

typedef struct vpf_gemm_epilogue {
  int mode;               /* VPF_EPI_* */
  int out_f32;            /* VPF_EPI_STORE: 1 = fp32 output, 0 = bf16 */
  int ldc;                /* row stride (elements) of out / out2 / resid / out_bf16 */
  int act;                /* VPF_ACT_* applied after bias */
  int aux_mode;           /* VPF_AUX_* applied after act */
  int ld_aux;
  int rg_shift, rg_ld;    /* row-group bias: rg_bias[(row >> rg_shift) * rg_ld + col] */
  unsigned int op_id;     /* dropout stream id */
  float alpha;            /* scale on the accumulator */
  float drop_p;           /* VPF_EPI_RESIDUAL: dropout probability (0 = off) */
  void *out;
  void *out2;             /* optional bf16 copy of (alpha*acc + biases) BEFORE act (saved for backward) */
  void *out_bf16;         /* VPF_EPI_RESIDUAL: optional bf16 copy of the result */
  const float *bias;      /* optional [N] */
  const float *rg_bias;   /* optional [(M >> rg_shift), rg_ld] */
  const void *aux;        /* optional bf16 [M, ld_aux] */
  const float *resid;     /* VPF_EPI_RESIDUAL: fp32 [M, ldc] */
  const unsigned long long *seed_ptr; /* device pointer to the step's dropout seed */
  /* fused group max (torch.max over the S points of a patch, utils.py:180,188) on the fp32 accumulators:
   * rows are grouped S at a time; gm_out_*[(row / S) * gm_ld + col] = max, gm_argmax = first maximal row.
   * `out` may be NULL when only the pooled result is wanted. */
  int gm_S, gm_ld;
  float *gm_out_f32;
  void *gm_out_bf16;
  unsigned char *gm_argmax;
  /* gm_cols = 1: the patch runs along the COLUMNS of C (call the GEMM transposed: A = weights [channels, K],
   * B = activations [points, K]); a thread then owns one channel and 32 consecutive points, the pool is pure
   * register work and gm_out_*[(col / S) * gm_ld + row] is written coalesced.  row_bias [M] is added after the max. */
  int gm_cols;
  const float *row_bias;
} vpf_gemm_epilogue;

int vpf_gemm_bf16(const void *A, int a_mn, int lda, const void *B, int b_mn, int ldb,
                  int M, int N, int K, int splits, const vpf_gemm_epilogue *epi,
                  void *stream);

/* ------------------------------------------------------------------ attention core
 * MultiHeadAttention.forward between the projections, partseg.py:71-84:
 * softmax(Q K^T * scale) -> dropout(p) -> . V, per (sample, head); head_dim = 64.
 * Q/K/V/O are bf16 token-row arrays ([B*L, ld], head h at columns h*64..h*64+63);
 * LSE fp32 [B*H, Lq] (log2 units) is saved for the backward pass. */
int vpf_attention_fwd(const void *Q, int ldq, const void *K, const void *V, int ldkv,
                      void *O, int ldo, float *LSE, int B, int H, int Lq, int Lk,
                      int head_dim, float scale, float drop_p,
                      const unsigned long long *seed_ptr, unsigned int op_id, void *stream);
/* delta_ws: fp32 scratch [B*H*Lq]. */
int vpf_attention_bwd(const void *Q, int ldq, const void *K, const void *V, int ldkv,
                      const void *O, int ldo, const void *dO, int lddo, const float *LSE,
                      float *delta_ws, void *dQ, int lddq, void *dK, void *dV, int lddkv,
                      int B, int H, int Lq, int Lk, int head_dim, float scale, float drop_p,
                      const unsigned long long *seed_ptr, unsigned int op_id, void *stream);

/* ------------------------------------------------------------- normalisation etc.
 * nn.LayerNorm (partseg.py:101-102,129,194; classifier.py:33), eps as given.
 * y_bf16 = LN(x + add[row % add_rows]) (optionally ReLU'd); xsum (optional) gets x + add. */
int vpf_layernorm_fwd(const void *x, int x_bf16, const float *add, int add_rows, float *xsum,
                      const float *gamma, const float *beta, void *y_bf16, float *mean,
                      float *rstd, int T, int D, float eps, int relu, void *stream);
/* dx = dres + LN'(dy) ; dpos[row % pos_rows] += dx ; dgamma/dbeta += ... (all optional but dx). */
int vpf_layernorm_bwd(const void *dy, int dy_bf16, const void *x, int x_bf16, const void *y_relu,
                      const float *mean, const float *rstd, const float *gamma, const float *dres,
                      void *dx, int dx_bf16, float *dgamma, float *dbeta, float *dpos, int pos_rows,
                      int T, int D, void *stream);
/* vpf_layernorm_bwd for (bf16 dy, fp32 x, fp32 dx, D % 128 == 0) that ALSO emits g_bf16 = dropout_mask(dx) * scale
 * (the byte-granular Residual mask of (seed, op_id), drop_p = 0 -> plain bf16 copy) and accumulates its column sums
 * into g_colsum (optional): the operand / bias gradient of the next block down the backward chain (partseg.py:208-213),
 * which otherwise costs a separate vpf_dropout_grad pass over dx. */
int vpf_layernorm_bwd_emit(const void *dy_bf16, const float *x, const float *mean, const float *rstd,
                           const float *gamma, const float *dres, float *dx, float *dgamma, float *dbeta,
                           float *dpos, int pos_rows, int T, int D, void *g_bf16, float *g_colsum, float drop_p,
                           const unsigned long long *seed_ptr, unsigned int op_id, void *stream);
/* Residual.dropout backward (partseg.py:201-213): out_bf16 = mask(g)/(1-p); colsum += column sums. */
int vpf_dropout_grad(const float *g, void *out_bf16, float *colsum, float p,
                     const unsigned long long *seed_ptr, unsigned int op_id, int T, int N, void *stream);
/* column sums (+=): sum/sumsq in fp64 (BatchNorm statistics), sum_f32 for bias gradients. */
int vpf_colsum(const void *x, int x_bf16, double *sum, double *sumsq, float *sum_f32,
               long long R, int C, void *stream);
/* nn.BatchNorm1d (utils.py:155,162; partseg.py:520,523): fold batch (training) or running (eval)
 * statistics into (scale, shift); updates running stats when training.  stats = [sum | sumsq]. */
int vpf_bn_finalize(const double *stats, long long R, const float *gamma, const float *beta,
                    float *running_mean, float *running_var, float momentum, float eps, int training,
                    float *scale, float *shift, float *mean, float *rstd, int C, void *stream);
int vpf_bn_apply(const void *x, int x_bf16, const float *scale, const float *shift, void *y,
                 int y_bf16, int relu, long long R, int C, void *stream);
/* train-mode BN (+ReLU) backward; red = fp64 scratch [3C]; dgamma/dbeta accumulate. */
int vpf_bn_bwd(const void *dy, int dy_bf16, const void *x, int x_bf16, const float *scale,
               const float *shift, const float *mean, const float *rstd, int relu, double *red,
               void *dx, int dx_bf16, float *dgamma, float *dbeta, long long R, int C, void *stream);
/* The same backward for an all-bf16 [R, 256] tensor fused with the row sum over groups of S consecutive rows of dx
 * (Group2Emb's split conv3 needs it, utils.py:183-185): gsum_bf16 / gsum_f32 [R/S, 256] (either may be null). */
int vpf_bn_bwd_gsum(const void *dy_bf16, const void *x_bf16, const float *scale, const float *shift,
                    const float *mean, const float *rstd, int relu, double *red, void *dx_bf16, float *dgamma,
                    float *dbeta, long long R, int C, int S, void *gsum_bf16, float *gsum_f32, void *stream);
/* nn.GELU (exact erf) of the MLP, partseg.py:196, as streaming kernels: h = gelu(z); dz = dh * gelu'(z) with the
 * column sums of dz (bias gradient) accumulated into colsum (optional). */
int vpf_gelu_fwd(const void *z_bf16, void *h_bf16, long long n, void *stream);
int vpf_gelu_bwd(const void *dh_bf16, const void *z_bf16, void *dz_bf16, float *colsum, long long R, int C, void *stream);
int vpf_cast_bf16(const float *x, void *y_bf16, long long n, void *stream);
int vpf_fill_zero(void *p, long long bytes, void *stream);

/* ------------------------------------------------------------------ pooling, thin ops
 * max over the S points of a group (utils.py:180,188) with argmax for the backward pass. */
int vpf_group_max_fwd(const void *x_bf16, void *out_bf16, float *out_f32, uint8_t *argmax,
                      int G, int S, int C, void *stream);
int vpf_group_max_bwd(const void *dout, int dout_bf16, const uint8_t *argmax, void *dx_bf16,
                      int accumulate, int G, int S, int C, void *stream);
/* sum over the S rows of each group (gradient of the broadcast in utils.py:183). */
int vpf_group_sum(const void *x_bf16, void *out_bf16, float *out_f32, int G, int S, int C, void *stream);
/* DropPath of the reference's Residual (partseg.py:201-213, timm DropPath: per-sample keep, survivors / (1 - p)):
 * scales[b] for one Residual from the device step seed and a site id; row_scale applies them to a [B*L, D] fp32 token
 * matrix (out may alias x).  Note the reference drops the WHOLE sum dropout(f(x)) + x, skip connection included. */
int vpf_droppath_scales(const unsigned long long *seed_ptr, unsigned int op_id, float p, int B, float *scales,
                        void *stream);
int vpf_row_scale(const float *x, const float *scales, int L, float *out, long long T, int D, void *stream);
/* out[n] += sum_k v[k] * W[k, n] for a bf16 [K, ldw] matrix window (fp32 accumulate): colsum(dY . W) = colsum(dY) . W,
 * the bias gradient of Group2Emb's first_conv.3 without a pass over the [B*G*S, 128] data gradient (utils.py:156). */
int vpf_vecmat_bf16(const float *v, const void *W_bf16, int ldw, int K, int N, float *out, void *stream);
/* cat(x.max(1)[0], x.mean(1)), partseg.py:547. */
int vpf_token_pool_fwd(const float *x, float *out, int *argmax, int B, int L, int D, void *stream);
int vpf_token_pool_bwd(const float *dout, const int *argmax, float *dx, int B, int L, int D, void *stream);
/* Linear(3, Co) / Conv1d(3, Co, 1) on xyz (classifier.py:32, partseg.py:499, utils.py:154). */
int vpf_linear3_fwd(const float *p, int ldp, const float *w, const float *b, const float *scale,
                    const float *shift, void *pre_bf16, void *act_bf16, int act, long long R, int Co,
                    void *stream);
int vpf_linear3_stats(const float *p, int ldp, const float *w, const float *b, double *stats,
                      long long R, int Co, void *stream);
int vpf_linear3_bwd(const void *dy, int dy_bf16, const float *p, int ldp, float *dW, float *db,
                    long long R, int Co, void *stream);
int vpf_linear3_bn_bwd(const void *dh_bf16, const float *p, int ldp, const float *w, const float *b,
                       const float *scale, const float *shift, const float *mean, const float *rstd,
                       double *red, float *dW, float *db, float *dgamma, float *dbeta, long long R,
                       int Co, void *stream);
/* Rearrange 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)', partseg.py:632 (fp32 NHWC, or NCHW
 * with nchw=1 which folds the permute of pretrain.py:179, -> bf16 rows). */
int vpf_patchify(const float *img, void *out_bf16, int B, int H, int W, int Ci, int P, int nchw, void *stream);
int vpf_add_scale(const float *a, const float *b, float *out, float alpha, long long n, void *stream);
/* Inference: fold an eval-mode BatchNorm1d (scale, shift from vpf_bn_finalize with training = 0) into the 1x1
 * convolution / Linear in front of it: Wout = bf16(scale[n] * W[n, :]), bout = scale * b + shift (utils.py:161-163). */
int vpf_bn_fold(const float *W, const float *b, const float *scale, const float *shift, void *Wout_bf16, float *bout,
                int N, int K, void *stream);
/* out[i] = x[i] * s[0], s a DEVICE scalar (upstream gradient of a scalar loss). */
int vpf_scale_by(const float *x, const float *s, float *out, long long n, void *stream);
/* dst[r, c] = alpha * src[r, c] over a [rows, cols] window of two row-strided fp32 arrays (padded head buffers). */
int vpf_copy2d(const float *src, int lds, float *dst, int ldd, long long rows, int cols, float alpha, void *stream);

/* ------------------------------------------------------------------ part segmentation
 * PointNetFeaturePropagation, utils.py:192-242: three nearest of the S centres for every point, ordered by (distance
 * ascending, index ascending), inverse-distance weights w = (1 / (d + 1e-8)) / sum. */
int vpf_three_nn(const float *pts, const float *centers, int B, int N, int S, int *idx, float *w, void *stream);
/* interpolated[b*N + n, :C] = sum_j w_j * feats[b*S + idx_j, :C] (utils.py:230), bf16, into a GEMM operand of row stride
 * ldo >= C + 8 whose columns C..C+2 receive the point coordinates (the cat of utils.py:233-234) and the rest zeros. */
int vpf_interp3_fwd(const void *feats_bf16, int ldf, const int *idx, const float *w, const float *pts, void *out_bf16,
                    int ldo, int B, int N, int S, int C, void *stream);
int vpf_interp3_bwd(const void *dout_bf16, int ldo, const int *idx, const float *w, float *dfeats, int ldd, int B, int N,
                    int S, int C, void *stream);
/* x.max(2), x.mean(2) over the groups (partseg.py:432-434) on a bf16 [B, L, ld] tensor; the backward ACCUMULATES. */
int vpf_token_pool_bf16_fwd(const void *x_bf16, int ld, float *out, int *argmax, int B, int L, int C, void *stream);
int vpf_token_pool_accum_bwd(const float *dout, const int *argmax, float *dx, int ld, int B, int L, int C, void *stream);
/* nn.LeakyReLU(slope), partseg.py:393. */
int vpf_leaky_relu_fwd(const float *x, float *y, float slope, long long n, void *stream);
int vpf_leaky_relu_bwd(const float *dy, const float *x, float *dx, float slope, long long n, void *stream);
/* first propagation Conv1d weight [Co, 3 + C] (xyz first) -> bf16 operand [Co, ldo] with the xyz columns moved behind the C
 * feature columns, and the inverse (+=) for its gradient. */
int vpf_permute_w(const float *W, void *Wout_bf16, int Co, int C, int ldo, void *stream);
int vpf_unpermute_dw(const float *dWp, float *dW, int Co, int C, int ldo, void *stream);
int vpf_copy2d_bf16(const void *src, int lds, void *dst, int ldd, long long rows, int cols, void *stream);

/* ------------------------------------------------------------------ loss + optimiser
 * NT-Xent (lightly 1.1.21 NTXentLoss; call sites pretrain.py:155,196,202). */
int vpf_l2norm_rows(const float *x, float *z, float *norm, int n, int D, void *stream);
/* Rows = this rank's 2*b_local normalised embeddings zr (out0 rows, then out1 rows); columns = n_c embeddings of all
 * ranks.  Local row i < b_local sits at column col_offset + i and its positive at col_offset + half + i (and vice versa
 * for i >= b_local).  Column j lives at row (j / zc_blk) * zc_ld + zc_base + j % zc_blk of zc (and of lse_all): with
 * every rank's PACKED block [imid rows | cmid rows] all-gathered once, zc_blk = 2b, zc_ld = 4b and zc_base selects the
 * loss term; a plain [n_c, D] buffer is zc_blk = zc_ld = n_c, zc_base = 0.
 * S_ws: fp32 scratch [n_r, n_c] -- holds the logits after fwd (keep it for bwd, which overwrites it); G_ws [n_r, D]. */
int vpf_ntxent_fwd(const float *zr, int n_r, const float *zc, int n_c, int D, int b_local,
                   int col_offset, int half, int zc_blk, int zc_ld, int zc_base, float temperature,
                   float *S_ws, float *lse_out, float *loss_out, void *stream);
int vpf_ntxent_bwd(const float *zr, const float *norm, int n_r, const float *zc, const float *lse_all,
                   int n_c, int D, int b_local, int col_offset, int half, int zc_blk, int zc_ld, int zc_base,
                   float temperature, float gscale, const float *upstream, float *S_ws, float *G_ws, float *dx,
                   void *stream);
/* Packed form: nseg (<= 4) loss terms that share zc / lse_all in ONE launch per stage.  zr, norm, lse_out, dx, G_ws hold
 * term t in rows [t * n_r, +n_r); S_ws is [nseg, n_r, n_c]; loss_out [nseg]; the column-map base of term t is
 * zc_base + t * zc_base_stride; gscale is a HOST array [nseg]. */
int vpf_ntxent_pack_fwd(const float *zr, int nseg, int n_r, const float *zc, int n_c, int D, int b_local,
                        int col_offset, int half, int zc_blk, int zc_ld, int zc_base, int zc_base_stride,
                        float temperature, float *S_ws, float *lse_out, float *loss_out, void *stream);
int vpf_ntxent_pack_bwd(const float *zr, const float *norm, int nseg, int n_r, const float *zc,
                        const float *lse_all, int n_c, int D, int b_local, int col_offset, int half, int zc_blk,
                        int zc_ld, int zc_base, int zc_base_stride, float temperature, const float *gscale,
                        const float *upstream, float *S_ws, float *G_ws, float *dx, void *stream);
/* nn.CrossEntropyLoss(label_smoothing = eps), mean reduction -- the fine-tune objective (ft_cls.py:145,176).
 * logits fp32 [n, ld] (C valid columns), labels int64 [n]; loss_out (+=, zero it first) and dlogits fp32 [n, ldd]
 * = d(mean loss)/d(logits), both produced in one pass. */
int vpf_ce_ls(const float *logits, int ld, const long long *labels, int n, int C, float eps, float *loss_out,
              float *dlogits, int ldd, void *stream);
/* torch.optim.AdamW step (pretrain.py:121-124,210) on a flat buffer, refreshing the bf16 shadow. */
int vpf_adamw(float *p, const float *g, float *m, float *v, void *shadow_bf16, long long n,
              const float *lr_ptr, float beta1, float beta2, float eps, float weight_decay,
              const long long *step_ptr, float grad_scale, void *stream);
/* state[0] += 1 (optimizer step), state[1] = next dropout seed. */
int vpf_step_advance(long long *state, void *stream);
/* n uniform indices in [0, N) drawn on the device from the step state's seed (state[1]) and a stream id: the FPS start
 * points of utils.py:71 (torch.randint there) without a host round trip or a library RNG kernel inside the captured
 * step.  out[i] = (hash(seed, op_id, i) * N) >> 32  -- restated in oracle/rng.py (draw_indices). */
int vpf_draw_indices(const long long *state, unsigned int op_id, int n, int N, long long *out, void *stream);

/* ---- scratch sizes (bytes) of the entries above that take caller-provided workspaces ---- */
size_t vpf_attention_bwd_workspace_bytes(int B, int H, int Lq);           /* delta_ws            */
size_t vpf_bn_bwd_workspace_bytes(int C);                                  /* red (vpf_bn_bwd)    */
size_t vpf_linear3_bn_bwd_workspace_bytes(int Co);                         /* red                 */
size_t vpf_ntxent_logits_workspace_bytes(int n_r, int n_c);                /* S_ws                */
size_t vpf_ntxent_grad_workspace_bytes(int n_r, int D);                    /* G_ws                */

#ifdef __cplusplus
}
#endif
#endif /* VPF_H_ */
