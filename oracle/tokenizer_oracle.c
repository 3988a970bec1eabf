/*
 * oracle/tokenizer_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the ViPFormer point-cloud
 * tokenizer (farthest-point sampling, kNN grouping, patch gather).  It exists
 * so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check / time the CUDA path against it.  Nothing under vipformer_b200/ may
 * import, link or call it.
 *
 * Each function cites the reference lines it restates
 * (paths relative to the upstream repo root):
 *   vipformer/model/pointcloud/utils.py:56-85    farthest_point_sample
 *   vipformer/model/pointcloud/utils.py:88-104   index_points
 *   vipformer/model/pointcloud/utils.py:107-119  knn_point
 *   vipformer/model/pointcloud/utils.py:122-141  square_distance
 *   vipformer/model/pointcloud/utils.py:6-38     divide_patches
 *
 * Pinned arithmetic (verified bit-for-bit against the reference's torch-CPU
 * output by tests/make_golden.py; see tests/golden/README.md):
 *   FPS   dist(p,c)  = ((dx*dx + dy*dy) + dz*dz), every op rounded to fp32,
 *                      NO fused multiply-add (utils.py:79 `sum((p-c)**2,-1)`).
 *         running    = min(running, dist), init 1e10f (utils.py:69,81)
 *         next       = argmax(running), LOWEST index on ties (utils.py:83,
 *                      torch.max returns the first maximal element on CPU).
 *   kNN   dot        = fmaf(cz,pz, fmaf(cy,py, cx*px))     (utils.py:138; what
 *                      the reference's K=3 sgemm evaluates to)
 *         d          = ((-2*dot) + |c|^2) + |p|^2            (utils.py:138-140)
 *         |v|^2      = ((x*x + y*y) + z*z) unfused           (utils.py:139-140)
 *         selection  = the `nsample` smallest by (d ascending, index
 *                      ascending), emitted in that order.  This is a legal
 *                      refinement of torch.topk(largest=False, sorted=False)
 *                      (utils.py:117) whose order is unspecified.
 *   gather           = neighbours[b,g,s,:] = pts[b, idx[b,g,s], :]; then the
 *                      reference's slice quirk (utils.py:36): for neighbour
 *                      SLOTS s in {0,1,2} only, every channel has the centre
 *                      subtracted; slots >= 3 stay in absolute coordinates.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VPF_ORACLE_OK 0
#define VPF_ORACLE_EINVAL -1
#define VPF_ORACLE_ENOMEM -2

/* utils.py:56-85.  pts [B,N,C] fp32 row-major (only channels 0..2 used),
 * start_idx [B] (the reference draws it with torch.randint, utils.py:71; here
 * it is an explicit input), out_idx [B,npoint] int64. */
int vpf_oracle_fps(const float *pts, int B, int N, int C, int npoint,
                   const int64_t *start_idx, int64_t *out_idx) {
  if (!pts || !start_idx || !out_idx || B < 0 || N <= 0 || C < 3 || npoint < 0)
    return VPF_ORACLE_EINVAL;
  float *running = (float *)malloc(sizeof(float) * (size_t)N);
  if (!running) return VPF_ORACLE_ENOMEM;
  for (int b = 0; b < B; ++b) {
    const float *p = pts + (size_t)b * N * C;
    int64_t far = start_idx[b];
    if (far < 0 || far >= N) { free(running); return VPF_ORACLE_EINVAL; }
    for (int i = 0; i < N; ++i) running[i] = 1e10f;
    for (int it = 0; it < npoint; ++it) {
      out_idx[(size_t)b * npoint + it] = far;
      const float cx = p[far * C + 0], cy = p[far * C + 1], cz = p[far * C + 2];
      float best = -INFINITY;
      int64_t besti = 0;
      for (int i = 0; i < N; ++i) {
        const float dx = p[(size_t)i * C + 0] - cx;
        const float dy = p[(size_t)i * C + 1] - cy;
        const float dz = p[(size_t)i * C + 2] - cz;
        float d = dx * dx;
        d = d + dy * dy;
        d = d + dz * dz;
        const float r = running[i] < d ? running[i] : d; /* torch.min */
        running[i] = r;
        if (r > best) { best = r; besti = i; } /* strict > keeps lowest index */
      }
      far = besti;
    }
  }
  free(running);
  return VPF_ORACLE_OK;
}

/* utils.py:88-104.  out[b,s,:] = points[b, idx[b,s], :]. */
int vpf_oracle_index_points(const float *points, int B, int N, int C,
                            const int64_t *idx, int S, float *out) {
  if (!points || !idx || !out) return VPF_ORACLE_EINVAL;
  for (int b = 0; b < B; ++b)
    for (int s = 0; s < S; ++s) {
      const int64_t j = idx[(size_t)b * S + s];
      if (j < 0 || j >= N) return VPF_ORACLE_EINVAL;
      memcpy(out + ((size_t)b * S + s) * C, points + ((size_t)b * N + j) * C,
             sizeof(float) * (size_t)C);
    }
  return VPF_ORACLE_OK;
}

static inline float sqnorm3(const float *v) {
  float s = v[0] * v[0];
  s = s + v[1] * v[1];
  s = s + v[2] * v[2];
  return s;
}

/* utils.py:122-141 with src = queries [B,S,Cq], dst = xyz [B,N,C]; only the
 * first three channels enter (divide_patches passes [:, :, :3], utils.py:19).
 * out [B,S,N]. */
int vpf_oracle_square_distance(const float *src, int B, int S, int Cq,
                               const float *dst, int N, int C, float *out) {
  if (!src || !dst || !out || C < 3 || Cq < 3) return VPF_ORACLE_EINVAL;
  for (int b = 0; b < B; ++b)
    for (int s = 0; s < S; ++s) {
      const float *c = src + ((size_t)b * S + s) * Cq;
      const float c2 = sqnorm3(c);
      for (int i = 0; i < N; ++i) {
        const float *p = dst + ((size_t)b * N + i) * C;
        const float dot = fmaf(c[2], p[2], fmaf(c[1], p[1], c[0] * p[0]));
        float d = -2.0f * dot;
        d = d + c2;
        d = d + sqnorm3(p);
        out[((size_t)b * S + s) * N + i] = d;
      }
    }
  return VPF_ORACLE_OK;
}

typedef struct { float d; int32_t i; } cand_t;

static int cand_less(const cand_t *a, const cand_t *b) {
  if (a->d < b->d) return 1;
  if (a->d > b->d) return 0;
  return a->i < b->i;
}

/* utils.py:107-119 under the stated order rule.  out_idx [B,S,nsample]. */
int vpf_oracle_knn(int nsample, const float *xyz, int B, int N, int C,
                   const float *new_xyz, int S, int Cq, int64_t *out_idx) {
  if (!xyz || !new_xyz || !out_idx || C < 3 || Cq < 3 || nsample <= 0 ||
      nsample > N)
    return VPF_ORACLE_EINVAL;
  cand_t *top = (cand_t *)malloc(sizeof(cand_t) * (size_t)nsample);
  float *p2 = (float *)malloc(sizeof(float) * (size_t)N);
  if (!top || !p2) { free(top); free(p2); return VPF_ORACLE_ENOMEM; }
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < N; ++i) p2[i] = sqnorm3(xyz + ((size_t)b * N + i) * C);
    for (int s = 0; s < S; ++s) {
      const float *c = new_xyz + ((size_t)b * S + s) * Cq;
      const float c2 = sqnorm3(c);
      int cnt = 0;
      for (int i = 0; i < N; ++i) {
        const float *p = xyz + ((size_t)b * N + i) * C;
        const float dot = fmaf(c[2], p[2], fmaf(c[1], p[1], c[0] * p[0]));
        float d = -2.0f * dot;
        d = d + c2;
        d = d + p2[i];
        cand_t k = {d, i};
        if (cnt == nsample && !cand_less(&k, &top[cnt - 1])) continue;
        /* insertion into the sorted prefix */
        int pos = cnt < nsample ? cnt : nsample - 1;
        while (pos > 0 && cand_less(&k, &top[pos - 1])) {
          top[pos] = top[pos - 1];
          --pos;
        }
        top[pos] = k;
        if (cnt < nsample) ++cnt;
      }
      for (int j = 0; j < nsample; ++j)
        out_idx[((size_t)b * S + s) * nsample + j] = top[j].i;
    }
  }
  free(top);
  free(p2);
  return VPF_ORACLE_OK;
}

/* utils.py:6-38.  neighbors [B,G,S,C], centers [B,G,C]; fps_idx [B,G] and
 * knn_idx [B,G,S] are optional (may be NULL). */
int vpf_oracle_divide_patches(const float *pts, int B, int N, int C, int G,
                              int S, const int64_t *start_idx,
                              float *neighbors, float *centers,
                              int64_t *fps_idx, int64_t *knn_idx) {
  if (!pts || !neighbors || !centers) return VPF_ORACLE_EINVAL;
  int rc;
  int64_t *fi = fps_idx ? fps_idx : (int64_t *)malloc(sizeof(int64_t) * (size_t)B * G);
  int64_t *ki = knn_idx ? knn_idx : (int64_t *)malloc(sizeof(int64_t) * (size_t)B * G * S);
  if (!fi || !ki) { rc = VPF_ORACLE_ENOMEM; goto done; }
  if ((rc = vpf_oracle_fps(pts, B, N, C, G, start_idx, fi))) goto done;
  if ((rc = vpf_oracle_index_points(pts, B, N, C, fi, G, centers))) goto done;
  if ((rc = vpf_oracle_knn(S, pts, B, N, C, centers, G, C, ki))) goto done;
  for (int b = 0; b < B; ++b)
    for (int g = 0; g < G; ++g) {
      const float *cen = centers + ((size_t)b * G + g) * C;
      for (int s = 0; s < S; ++s) {
        const int64_t j = ki[((size_t)b * G + g) * S + s];
        float *o = neighbors + (((size_t)b * G + g) * S + s) * C;
        const float *p = pts + ((size_t)b * N + j) * C;
        for (int ch = 0; ch < C; ++ch)
          o[ch] = (s < 3) ? p[ch] - cen[ch] : p[ch]; /* utils.py:36 quirk */
      }
    }
done:
  if (!fps_idx) free(fi);
  if (!knn_idx) free(ki);
  return rc;
}
