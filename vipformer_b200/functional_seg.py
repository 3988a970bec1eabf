"""Part-segmentation head of CrossFormer_partseg (vipformer/model/pointcloud/partseg.py:420-468) as an explicit kernel
sequence: everything after the encoder's per-layer taps -- shared LayerNorm, channel concat, group max / mean, label
embedding, PointNetFeaturePropagation (3-NN inverse-distance interpolation + 2 x {Conv1d, BN, ReLU}, utils.py:192-242) and the
per-point classifier conv1/bn1/dp1/conv2/bn2/conv3.

Layout: rows x channels throughout (the reference's [B, C, N] tensors transposed), so its transposes / permutes vanish.  Two
reference concatenations never materialise: cat([points1, interpolated]) becomes 8 extra K columns of the interpolation
output (xyz + zero padding; the first propagation weight is column-permuted to match), and the per-sample global feature
repeated over the N points (x_max, x_avg, label embedding: partseg.py:436-445) enters conv1 as a per-sample bias
u = W1[:, 1024:] . g + b1 through the GEMM's row-group-bias epilogue -- which needs N to be a power of two.
"""
import math
from types import SimpleNamespace as NS

import torch

from . import ops
from .functional import _dgrad, _empty, _pad8, _wgrad
from .ops import BF16, EPI_ATOMIC_ADD, F32


def _bias_grad(g, gfull, cout, Gb, like):
    if cout % 8 == 0:
        ops.colsum(g, sum32=Gb)
    else:
        tmp = ops.zeros_(_empty((_pad8(cout),), F32, like))
        ops.colsum(gfull, sum32=tmp)
        ops.add_scale(tmp[:cout], Gb, 1.0, out=Gb)


def partseg_head_fwd(taps, pts, ctr, onehot, W, bn, cfg, training, seed, op_id, save=True):
    """taps: k fp32 [B*G, D]; pts fp32 [B,N,3]; ctr fp32 [B,G,3]; onehot fp32 [B,16] -> logits fp32 view [B*N, P]."""
    B, N, G, D, k, P = cfg.B, cfg.N, cfg.G, cfg.D, len(taps), cfg.P
    if N & (N - 1):
        raise NotImplementedError("CrossFormer_partseg kernels need a power-of-two number of points (row-group bias epilogue)")
    kD, R = k * D, B * N
    # shared LayerNorm on every tap (partseg.py:422), concatenated along channels (:424-428)
    X = _empty((B * G, kD), BF16, pts)
    ln = []
    for i, t in enumerate(taps):
        Ti, mean, rstd, _ = ops.layernorm_fwd(t, W.ln_w, W.ln_b)
        ops.copy2d_bf16(Ti, X[:, i * D:(i + 1) * D])
        ln.append((t, mean, rstd))
    pooled, am = ops.token_pool_bf16_fwd(X, B, G, kD)                      # [B, 2kD] = (x_max || x_avg), :430-434
    # label embedding: Conv1d(16, 64, no bias) -> BN1d -> LeakyReLU(0.2) on the one-hot [B, 16, 1] (:391-393, :441-443)
    oh = ops.cast_bf16(onehot.float().contiguous())
    ylc = _empty((B, 64), F32, pts)
    ops.gemm(oh, W.lc_w, ylc)
    zlc, st_lc = ops.bn_forward(ylc, W.lc_bn_w, W.lc_bn_b, bn.lc_rm, bn.lc_rv, training, False, out_dtype=F32)
    clsf = ops.leaky_relu_fwd(zlc, 0.2)
    gfeat = _empty((B, 2 * kD + 64), F32, pts)
    ops.copy2d(pooled, gfeat[:, :2 * kD])
    ops.copy2d(clsf, gfeat[:, 2 * kD:])
    gfeat_bf = ops.cast_bf16(gfeat)
    u = _empty((B, 512), F32, pts)
    ops.gemm(gfeat_bf, W.c1_w[:, 1024:], u, bias=W.c1_b)                   # conv1's share of the repeated global feature
    # PointNetFeaturePropagation (utils.py:209-242)
    idx, w3 = ops.three_nn(pts, ctr)
    Kp = kD + 8
    F0 = ops.interp3_fwd(X, idx, w3, pts, Kp)                              # [R, Kp] = (interpolated || xyz || 0)
    p1w = ops.permute_w(W.p1_w_f32, kD, Kp)
    Cp1 = p1w.shape[0]
    y_p1 = _empty((R, Cp1), BF16, pts)
    ops.gemm(F0, p1w, y_p1, bias=W.p1_b)
    a_p1, st_p1 = ops.bn_forward(y_p1, W.p1_bn_w, W.p1_bn_b, bn.p1_rm, bn.p1_rv, training, True)
    y_p2 = _empty((R, 1024), BF16, pts)
    ops.gemm(a_p1, W.p2_w, y_p2, bias=W.p2_b)
    f0, st_p2 = ops.bn_forward(y_p2, W.p2_bn_w, W.p2_bn_b, bn.p2_rm, bn.p2_rv, training, True)
    # per-point classifier (partseg.py:452-463)
    y1 = _empty((R, 512), BF16, pts)
    ops.gemm(f0, W.c1_w[:, :1024], y1, rg_bias=u, rg_shift=int(math.log2(N)))
    a1f, st1 = ops.bn_forward(y1, W.bn1_w, W.bn1_b, bn.rm1, bn.rv1, training, True, out_dtype=F32)
    p_drop = float(getattr(cfg, 'p_dp1', 0.5)) if training else 0.0
    a1 = ops.dropout_grad(a1f, p_drop, seed, op_id)                        # dp1 (forward mask == backward mask by construction)
    y2 = _empty((R, 256), BF16, pts)
    ops.gemm(a1, W.c2_w, y2, bias=W.c2_b)
    a2, st2 = ops.bn_forward(y2, W.bn2_w, W.bn2_b, bn.rm2, bn.rv2, training, True)
    lbuf = _empty((R, _pad8(P)), F32, pts)
    logits = lbuf[:, :P]
    ops.gemm(a2, W.c3_w, logits, bias=W.c3_b)
    ctx = None
    if save:
        ctx = NS(ln=ln, X=X, am=am, oh=oh, ylc=ylc, zlc=zlc, st_lc=st_lc, gfeat_bf=gfeat_bf, idx=idx, w3=w3, F0=F0, p1w=p1w,
                 y_p1=y_p1, a_p1=a_p1, st_p1=st_p1, y_p2=y_p2, f0=f0, st_p2=st_p2, y1=y1, st1=st1, a1=a1, y2=y2, a2=a2, st2=st2,
                 p_drop=p_drop, Kp=Kp)
    return logits, ctx


def partseg_head_bwd(dlogits, c, W, G, cfg, seed, op_id):
    """dlogits fp32 [B*N, P] -> list of k fp32 [B*G, D] tap gradients; parameter gradients accumulate into G."""
    B, N, Gn, D, P = cfg.B, cfg.N, cfg.G, cfg.D, cfg.P
    k = len(c.ln)
    kD, R = k * D, B * N
    like = dlogits
    # conv3
    gp = ops.zeros_(_empty((R, _pad8(P)), F32, like)) if P % 8 else _empty((R, P), F32, like)
    ops.copy2d(dlogits, gp[:, :P])
    g3full = ops.dropout_grad(gp, 0.0, None, 0)
    g3 = g3full[:, :P]
    _bias_grad(g3, g3full, P, G.c3_b, like)
    _wgrad(g3, c.a2, G.c3_w)
    da2 = _empty((R, 256), BF16, like)
    _dgrad(g3, W.c3_w, da2)
    # bn2 + conv2 + dp1
    dy2 = ops.bn_backward(da2, c.y2, c.st2, True, G.bn2_w, G.bn2_b)
    ops.colsum(dy2, sum32=G.c2_b)
    _wgrad(dy2, c.a1, G.c2_w)
    da1f = _empty((R, 512), F32, like)
    _dgrad(dy2, W.c2_w, da1f)
    da1 = ops.dropout_grad(da1f, c.p_drop, seed, op_id)
    # bn1 + conv1 (per-point part and per-sample bias part)
    dy1 = ops.bn_backward(da1, c.y1, c.st1, True, G.bn1_w, G.bn1_b)
    full = W.c1_w.shape[1]
    _wgrad(dy1, c.f0, G.c1_w[:, :1024], ldc=full)
    df0 = _empty((R, 1024), BF16, like)
    _dgrad(dy1, W.c1_w[:, :1024], df0)
    du_bf, du = ops.group_sum(dy1, B, N, 512, want_bf16=True, want_f32=True)
    ops.colsum(du, sum32=G.c1_b)
    _wgrad(du_bf, c.gfeat_bf, G.c1_w[:, 1024:], ldc=full)
    dgfeat = _empty((B, 2 * kD + 64), F32, like)
    _dgrad(du_bf, W.c1_w[:, 1024:], dgfeat)
    # propagation MLP
    dy_p2 = ops.bn_backward(df0, c.y_p2, c.st_p2, True, G.p2_bn_w, G.p2_bn_b)
    ops.colsum(dy_p2, sum32=G.p2_b)
    _wgrad(dy_p2, c.a_p1, G.p2_w)
    Cp1 = c.p1w.shape[0]
    da_p1 = _empty((R, Cp1), BF16, like)
    _dgrad(dy_p2, W.p2_w, da_p1)
    dy_p1 = ops.bn_backward(da_p1, c.y_p1, c.st_p1, True, G.p1_bn_w, G.p1_bn_b)
    ops.colsum(dy_p1, sum32=G.p1_b)
    dWp = ops.zeros_(_empty((Cp1, c.Kp), F32, like))
    _wgrad(dy_p1, c.F0, dWp)
    ops.unpermute_dw(dWp, G.p1_w, kD)
    dF0 = _empty((R, c.Kp), BF16, like)
    _dgrad(dy_p1, c.p1w, dF0)
    # back to the group features: interpolation + group max / mean
    dX = ops.zeros_(_empty((B * Gn, kD), F32, like))
    ops.interp3_bwd(dF0, c.idx, c.w3, dX, B, N, Gn, kD)
    dpooled = _empty((B, 2 * kD), F32, like)
    ops.copy2d(dgfeat[:, :2 * kD], dpooled)
    ops.token_pool_accum_bwd(dpooled, c.am, dX, B, Gn, kD)
    # label embedding
    dclsf = _empty((B, 64), F32, like)
    ops.copy2d(dgfeat[:, 2 * kD:], dclsf)
    dz = ops.leaky_relu_bwd(dclsf, c.zlc, 0.2)
    dylc = ops.bn_backward(dz, c.ylc, c.st_lc, False, G.lc_bn_w, G.lc_bn_b, out_dtype=BF16)
    _wgrad(dylc, c.oh, G.lc_w)
    # shared LayerNorm of every tap
    dtaps = []
    for i, (t, mean, rstd) in enumerate(c.ln):
        dTi = _empty((B * Gn, D), F32, like)
        ops.copy2d(dX[:, i * D:(i + 1) * D], dTi)
        dtaps.append(ops.layernorm_bwd(dTi, t, mean, rstd, W.ln_w, dgamma=G.ln_w, dbeta=G.ln_b))
    return dtaps
