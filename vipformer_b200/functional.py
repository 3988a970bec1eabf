"""Forward/backward of the ViPFormer blocks as explicit kernel sequences (no autograd inside).

Every function takes plain tensors and returns (outputs, ctx); the matching *_bwd consumes ctx,
accumulates parameter gradients IN PLACE into the fp32 grad tensors it is handed (the kernels
use red.global.add / atomics, so handing in views of one flat gradient buffer is free) and
returns the input gradients.  `modules.py` wraps these in torch.autograd.Function.

Storage conventions: residual stream and small per-sample tensors fp32; GEMM operands and the big
per-point / per-group-point activations bf16; all accumulation fp32 (fp64 for BatchNorm sums).
Reference: vipformer/model/pointcloud/partseg.py:15-342,473-680, utils.py:144-189, classifier.py:25-50.
"""
import math
from types import SimpleNamespace as NS

import torch

from . import ops
from .ops import ACT_GELU, ACT_RELU, AUX_GELU_GRAD, BF16, EPI_ATOMIC_ADD, EPI_RESIDUAL, F32


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _wgrad(dy, x, dW, ldc=None):
    """dW[N_out, K_in] += dy[T, N_out]^T @ x[T, K_in]   (both operands MN-major, split-K, atomic accumulate)."""
    ops.gemm(dy, x, dW, a_mn=True, b_mn=True, mode=EPI_ATOMIC_ADD, ldc=ldc)


def _dgrad(dy, w, out, **kw):
    """out[T, K_in] = dy[T, N_out] @ w[N_out, K_in]   (B consumed MN-major: no transposed weight copy)."""
    return ops.gemm(dy, w, out, b_mn=True, **kw)


_EMIT = __import__("os").environ.get("VPF_LN_EMIT", "1") != "0"   # measurement switch: LayerNorm backward emits the next operand


# =============================================================================== MLP (partseg.py:191-198)
def _mlp_fwd(x1, W, T, D, p_drop, seed, op_id, save):
    xn2, mean2, rstd2, _ = ops.layernorm_fwd(x1, W.ln2_w, W.ln2_b)
    F_ = W.w1.shape[0]
    # GELU runs as a streaming kernel at full occupancy (measured faster than in the 16-warp GEMM epilogue)
    z = _empty((T, F_), BF16, x1)
    ops.gemm(xn2, W.w1, z, bias=W.b1)
    h = ops.gelu_fwd(z)
    if not save:
        z = None
    x2 = _empty((T, D), F32, x1)
    ops.gemm(h, W.w2, x2, bias=W.b2, mode=EPI_RESIDUAL, resid=x1, drop_p=p_drop, seed=seed, op_id=op_id)
    return x2, NS(xn2=xn2, mean2=mean2, rstd2=rstd2, h=h, z=z, x1=x1)


def _ln_bwd(dy, x, mean, rstd, gamma, emit, **kw):
    """LayerNorm backward; with `emit` = (drop_p, seed, op_id, colsum) also the masked bf16 copy of the result that the next
    block down the chain consumes (ops.layernorm_bwd_emit) -> (dx, g or None)."""
    if emit is not None and _EMIT and ops.can_emit(dy, x, x.shape[1]):
        return ops.layernorm_bwd_emit(dy, x, mean, rstd, gamma, emit, **kw)
    return ops.layernorm_bwd(dy, x, mean, rstd, gamma, **kw), None


def _mlp_bwd(dx2, c, W, G, T, D, p_drop, seed, op_id, g2=None, emit=None):
    """g2: the masked bf16 copy of dx2 when the block above already produced it; emit: what to produce for the block below."""
    if g2 is None:
        g2 = ops.dropout_grad(dx2, p_drop, seed, op_id, colsum=G.b2)
    _wgrad(g2, c.h, G.w2)
    F_ = W.w1.shape[0]
    dh = _empty((T, F_), BF16, dx2)
    _dgrad(g2, W.w2, dh)
    dz = ops.gelu_bwd(dh, c.z, colsum=G.b1)
    _wgrad(dz, c.xn2, G.w1)
    dxn2 = _empty((T, D), BF16, dx2)      # gradient w.r.t. the LayerNorm output: a bf16 operand like dh / dz / dqkv
    _dgrad(dz, W.w1, dxn2)
    return _ln_bwd(dxn2, c.x1, c.mean2, c.rstd2, W.ln2_w, emit, dres=dx2, dgamma=G.ln2_w, dbeta=G.ln2_b)


# ============================================================ SelfAttentionLayer (partseg.py:170-188)
def sa_layer_fwd(x_prev, pos, W, cfg, seed, op_base, save=True):
    """x_prev fp32 [B*L, D]; pos fp32 [pos_rows, D] (re-added before every layer, Encoder.forward :331-335)."""
    B, L, D, H = cfg.B, cfg.L, cfg.D, cfg.H
    T = B * L
    xn, mean1, rstd1, xin = ops.layernorm_fwd(x_prev, W.ln1_w, W.ln1_b, add=pos, want_xsum=True)
    qkv = _empty((T, 3 * D), BF16, x_prev)
    ops.gemm(xn, W.wqkv, qkv)
    o, lse = ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, L, L, cfg.scale, cfg.p_attn, seed, op_base)
    x1 = _empty((T, D), F32, x_prev)
    ops.gemm(o, W.wo, x1, bias=W.bo, mode=EPI_RESIDUAL, resid=xin, drop_p=cfg.p_res1, seed=seed, op_id=op_base + 1)
    s1 = s2 = None
    p_path = getattr(cfg, "p_path", 0.0)
    if p_path > 0.0:      # DropPath of the attention Residual: the WHOLE sum dropout(f(x)) + x is scaled per sample (partseg.py:212)
        s1 = ops.droppath_scales(seed, op_base + 3, p_path, B, x1)
        ops.row_scale(x1, s1, L, out=x1)
    x2, cm = _mlp_fwd(x1, W, T, D, cfg.p_res2, seed, op_base + 2, save)
    if p_path > 0.0:      # ... and of the MLP Residual, an independent draw
        s2 = ops.droppath_scales(seed, op_base + 4, p_path, B, x2)
        ops.row_scale(x2, s2, L, out=x2)
    ctx = NS(xn=xn, mean1=mean1, rstd1=rstd1, xin=xin, qkv=qkv, o=o, lse=lse, mlp=cm, s1=s1, s2=s2) if save else None
    return x2, ctx


def sa_layer_bwd(dx2, c, W, G, cfg, seed, op_base, dpos, g2=None, emit=None):
    """-> (dx, g): g = masked bf16 copy of dx for the layer below when `emit` = (drop_p, seed, op_id, colsum) asks for it.
    The LayerNorm backward of the MLP block emits this layer's own attention-residual operand (g1) the same way."""
    B, L, D, H = cfg.B, cfg.L, cfg.D, cfg.H
    T = B * L
    own_emit = (cfg.p_res1, seed, op_base + 1, G.bo)
    if c.s2 is not None:      # DropPath: the gradient of the whole Residual output is scaled per sample before it splits
        dx2, g2, own_emit = ops.row_scale(dx2, c.s2, L), None, None
    dx1, g1 = _mlp_bwd(dx2, c.mlp, W, G, T, D, cfg.p_res2, seed, op_base + 2, g2=g2, emit=own_emit)
    if c.s1 is not None:
        ops.row_scale(dx1, c.s1, L, out=dx1)
    if g1 is None:
        g1 = ops.dropout_grad(dx1, cfg.p_res1, seed, op_base + 1, colsum=G.bo)
    _wgrad(g1, c.o, G.wo)
    do = _empty((T, D), BF16, dx2)
    _dgrad(g1, W.wo, do)
    dqkv = _empty((T, 3 * D), BF16, dx2)
    qkv = c.qkv
    ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], c.o, do, c.lse, dqkv[:, :D], dqkv[:, D:2 * D],
                      dqkv[:, 2 * D:], B, H, L, L, cfg.scale, cfg.p_attn, seed, op_base)
    _wgrad(dqkv, c.xn, G.wqkv)
    dxn = _empty((T, D), BF16, dx2)
    _dgrad(dqkv, W.wqkv, dxn)
    return _ln_bwd(dxn, c.xin, c.mean1, c.rstd1, W.ln1_w, emit, dres=dx1, dgamma=G.ln1_w, dbeta=G.ln1_b, dpos=dpos)


# =========================================================== CrossAttentionLayer (partseg.py:144-167)
def ca_layer_fwd(xq_prev, pos, kv_in, W, cfg, seed, op_base, save=True):
    """xq_prev fp32 [B*L, D] (+pos); kv_in bf16 or fp32 [B*Lk, D] (no positional term, Encoder.forward :326)."""
    B, L, Lk, D, H = cfg.B, cfg.L, cfg.Lk, cfg.D, cfg.H
    T, Tk = B * L, B * Lk
    qn, meanq, rstdq, xin = ops.layernorm_fwd(xq_prev, W.qn_w, W.qn_b, add=pos, want_xsum=True)
    kvn, meank, rstdk, _ = ops.layernorm_fwd(kv_in, W.kvn_w, W.kvn_b)
    q = _empty((T, D), BF16, xq_prev)
    ops.gemm(qn, W.wq, q)
    kvp = _empty((Tk, 2 * D), BF16, xq_prev)
    ops.gemm(kvn, W.wkv, kvp)
    o, lse = ops.attention_fwd(q, kvp[:, :D], kvp[:, D:], B, H, L, Lk, cfg.scale, cfg.p_attn, seed, op_base)
    x1 = _empty((T, D), F32, xq_prev)
    ops.gemm(o, W.wo, x1, bias=W.bo, mode=EPI_RESIDUAL, resid=xin, drop_p=cfg.p_res1, seed=seed, op_id=op_base + 1)
    x2, cm = _mlp_fwd(x1, W, T, D, cfg.p_res2, seed, op_base + 2, save)
    ctx = NS(qn=qn, meanq=meanq, rstdq=rstdq, xin=xin, kvn=kvn, meank=meank, rstdk=rstdk, kv_in=kv_in, q=q, kvp=kvp,
             o=o, lse=lse, mlp=cm) if save else None
    return x2, ctx


def ca_layer_bwd(dx2, c, W, G, cfg, seed, op_base, dpos, need_dkv=True, g2=None):
    B, L, Lk, D, H = cfg.B, cfg.L, cfg.Lk, cfg.D, cfg.H
    T, Tk = B * L, B * Lk
    dx1, g1 = _mlp_bwd(dx2, c.mlp, W, G, T, D, cfg.p_res2, seed, op_base + 2, g2=g2,
                       emit=(cfg.p_res1, seed, op_base + 1, G.bo))
    if g1 is None:
        g1 = ops.dropout_grad(dx1, cfg.p_res1, seed, op_base + 1, colsum=G.bo)
    _wgrad(g1, c.o, G.wo)
    do = _empty((T, D), BF16, dx2)
    _dgrad(g1, W.wo, do)
    dq = _empty((T, D), BF16, dx2)
    dkvp = _empty((Tk, 2 * D), BF16, dx2)
    ops.attention_bwd(c.q, c.kvp[:, :D], c.kvp[:, D:], c.o, do, c.lse, dq, dkvp[:, :D], dkvp[:, D:], B, H, L, Lk,
                      cfg.scale, cfg.p_attn, seed, op_base)
    _wgrad(dq, c.qn, G.wq)
    dqn = _empty((T, D), BF16, dx2)
    _dgrad(dq, W.wq, dqn)
    dxq = ops.layernorm_bwd(dqn, c.xin, c.meanq, c.rstdq, W.qn_w, dres=dx1, dgamma=G.qn_w, dbeta=G.qn_b, dpos=dpos)
    _wgrad(dkvp, c.kvn, G.wkv)
    kv_bf16 = c.kv_in.dtype == BF16
    dkvn = _empty((Tk, D), BF16 if kv_bf16 else F32, dx2)
    _dgrad(dkvp, W.wkv, dkvn)
    dkv = ops.layernorm_bwd(dkvn, c.kv_in, c.meank, c.rstdk, W.kvn_w, dgamma=G.kvn_w, dbeta=G.kvn_b,
                            out_dtype=BF16 if kv_bf16 else F32)
    return dxq, dkv


# ==================================================================== Group2Emb (utils.py:144-189)
def group2emb_fwd(nb, W, bn, cfg, training, save=True):
    """nb fp32 [B,G,S,3] -> tokens fp32 [B*G, D].  bn holds the two BatchNorm1d running-stat pairs."""
    Gt, S, D = cfg.Gt, cfg.S, cfg.D
    R = Gt * S
    assert S & (S - 1) == 0, "group_size must be a power of two (row-group bias epilogue)"
    if training:
        stats1 = ops.linear3_stats(nb, 3, W.w1, W.b1, R)
    else:
        stats1 = None
    st1 = ops.bn_stats_finalize(stats1, R, W.bn1_w, W.bn1_b, bn.rm1, bn.rv1, training)
    _, h1 = ops.linear3_fwd(nb, 3, W.w1, W.b1, R, scale=st1.scale, shift=st1.shift, act=ACT_RELU)
    # conv2 with the per-patch max pooled from the fp32 accumulators in the GEMM epilogue (utils.py:180)
    f2 = _empty((R, 128), BF16, nb)
    gmax = _empty((Gt, 128), BF16, nb)
    am2 = _empty((Gt, 128), torch.uint8, nb)
    ops.gemm(h1, W.w2, f2, bias=W.b2, gm_S=S, gm_bf16=gmax, gm_argmax=am2)
    # conv3 on cat([global, local]) split into a per-group and a per-point half (saves 1/4 of the block's FLOPs)
    u = _empty((Gt, 256), F32, nb)
    if not training:
        # inference: the eval-mode BatchNorm (running statistics) is folded into conv3 -- scaled weight rows, folded bias,
        # ReLU in the GEMM epilogue -- so neither y3 nor a separate normalisation pass exists (utils.py:161-163)
        st3 = ops.bn_stats_finalize(None, R, W.bn3_w, W.bn3_b, bn.rm3, bn.rv3, False)
        w3s, b3s = ops.bn_fold(W.w3_f32, W.b3, st3.scale, st3.shift)
        ops.gemm(gmax, w3s[:, :128], u, bias=b3s)
        h3 = _empty((R, 256), BF16, nb)
        ops.gemm(f2, w3s[:, 128:], h3, rg_bias=u, rg_shift=int(math.log2(S)), act=ACT_RELU)
        y3 = None
    else:
        ops.gemm(gmax, W.w3[:, :128], u, bias=W.b3)
        y3 = _empty((R, 256), BF16, nb)
        ops.gemm(f2, W.w3[:, 128:], y3, rg_bias=u, rg_shift=int(math.log2(S)))
        h3, st3 = ops.bn_forward(y3, W.bn3_w, W.bn3_b, bn.rm3, bn.rv3, training, True)
    # conv4 + max over the patch (utils.py:187-188): pooled from the fp32 accumulators in the GEMM epilogue, the [R, D]
    # pre-pool tensor is never stored.  (The transposed, register-only variant `gm_cols=True` measured slower: 864 vs 600 us.)
    tok = _empty((Gt, D), F32, nb)
    am4 = _empty((Gt, D), torch.uint8, nb)
    ops.gemm(h3, W.w4, None, bias=W.b4, gm_S=S, gm_f32=tok, gm_argmax=am4)
    ctx = NS(nb=nb, st1=st1, h1=h1, f2=f2, gmax=gmax, am2=am2, y3=y3, st3=st3, h3=h3, am4=am4) if save else None
    return tok, ctx


def group2emb_bwd(dtok, c, W, G, cfg):
    Gt, S, D = cfg.Gt, cfg.S, cfg.D
    if not c.st1.training:
        raise NotImplementedError("backward through an eval-mode Group2Emb (running-statistics BatchNorm) is not built")
    R = Gt * S
    ops.colsum(dtok, sum32=G.b4)
    dy4 = ops.group_max_bwd(dtok, c.am4, Gt, S, D)
    _wgrad(dy4, c.h3, G.w4)
    dh3 = _empty((R, 256), BF16, dtok)
    _dgrad(dy4, W.w4, dh3)
    del dy4
    # BatchNorm backward and the per-patch row sum of its result (for the per-patch half of the split conv3) in one pass
    dy3, dug, _ = ops.bn_backward_gsum(dh3, c.y3, c.st3, True, G.bn3_w, G.bn3_b, S)
    del dh3
    _wgrad(dy3, c.f2, G.w3[:, 128:])
    df2 = _empty((R, 128), BF16, dtok)
    _dgrad(dy3, W.w3[:, 128:], df2)
    del dy3
    cs3 = ops.zeros_(_empty((256,), F32, dtok))
    ops.colsum(dug, sum32=cs3)                       # = column sums of dy3 over ALL rows (dug holds the per-patch sums)
    ops.add_scale(G.b3, cs3, 1.0, out=G.b3)
    _wgrad(dug, c.gmax, G.w3[:, :128])
    dgmax = _empty((Gt, 128), F32, dtok)
    _dgrad(dug, W.w3[:, :128], dgmax)
    ops.group_max_bwd(dgmax, c.am2, Gt, S, 128, dx=df2)
    # bias gradient of conv2 = colsum(df2) = colsum(dy3) . W3[:, 128:] + colsum(dgmax): two tiny kernels instead of a pass
    # over the 0.5 GB df2 (and without the bf16 rounding of its entries)
    ops.vecmat_bf16(cs3, W.w3[:, 128:], G.b2)
    ops.colsum(dgmax, sum32=G.b2)
    _wgrad(df2, c.h1, G.w2)
    dh1 = _empty((R, 64), BF16, dtok)
    _dgrad(df2, W.w2, dh1)
    ops.linear3_bn_bwd(dh1, c.nb, 3, W.w1, W.b1, c.st1, G.w1, G.b1, G.bn1_w, G.bn1_b, R)


# ====================================================== PointCloudInputAdapter (classifier.py:25-50)
def adapter_fwd(pts, W, save=True):
    """pts fp32 [B,N,C] -> bf16 [B*N, D]."""
    B, N, C = pts.shape
    R = B * N
    y1, _ = ops.linear3_fwd(pts, C, W.w1, W.b1, R, want_pre=True)
    h, mean, rstd, _ = ops.layernorm_fwd(y1, W.ln_w, W.ln_b, relu=True)
    e = _empty((R, W.w2.shape[0]), BF16, pts)
    ops.gemm(h, W.w2, e, bias=W.b2)
    return e, (NS(pts=pts, y1=y1, h=h, mean=mean, rstd=rstd) if save else None)


def adapter_bwd(de, c, W, G):
    B, N, C = c.pts.shape
    R = B * N
    if de.dtype != BF16:
        de = ops.dropout_grad(de, 0.0, None, 0)
    ops.colsum(de, sum32=G.b2)
    _wgrad(de, c.h, G.w2)
    dh = _empty((R, 64), BF16, de)
    _dgrad(de, W.w2, dh)
    dy1 = ops.layernorm_bwd(dh, c.y1, c.mean, c.rstd, W.ln_w, y_relu=c.h, dgamma=G.ln_w, dbeta=G.ln_b, out_dtype=BF16)
    ops.linear3_bwd(dy1, c.pts, C, G.w1, G.b1, R)


# ============================================================== position_emb (partseg.py:498-501)
def posemb_fwd(center, W, save=True):
    """center fp32 [B,G,3] -> fp32 [B*G, D]."""
    R = center.shape[0] * center.shape[1]
    z, h = ops.linear3_fwd(center, center.shape[2], W.w1, W.b1, R, want_pre=True, act=ACT_GELU)
    pos = _empty((R, W.w2.shape[0]), F32, center)
    ops.gemm(h, W.w2, pos, bias=W.b2)
    return pos, (NS(center=center, z=z, h=h) if save else None)


def posemb_bwd(dpos, c, W, G):
    R = c.z.shape[0]
    g = ops.dropout_grad(dpos, 0.0, None, 0, colsum=G.b2)
    _wgrad(g, c.h, G.w2)
    dz = _empty((R, 128), BF16, dpos)
    _dgrad(g, W.w2, dz, aux=c.z, aux_mode=AUX_GELU_GRAD)
    ops.linear3_bwd(dz, c.center, c.center.shape[2], G.w1, G.b1, R)


# ================================================================= patch2emb (partseg.py:631-634)
def patch2emb_fwd(imgs, W, patch, save=True, nchw=False):
    """imgs fp32 [B,H,W,3] NHWC (or [B,3,H,W] with nchw) -> fp32 [B*np, D]."""
    P = ops.patchify(imgs, patch, nchw)
    e = _empty((P.shape[0], W.w.shape[0]), F32, imgs)
    ops.gemm(P, W.w, e, bias=W.b)
    return e, (NS(P=P) if save else None)


def patch2emb_bwd(de, c, G):
    g = ops.dropout_grad(de, 0.0, None, 0, colsum=G.b)
    _wgrad(g, c.P, G.w)


# ===================================================== pooling + BN/ReLU/Linear heads
# latent_head (partseg.py:519-525: 2 x {BN1d, ReLU, Linear no bias}) and finetune_head (partseg.py:573-582: 3 x {BN1d, ReLU,
# Linear with bias}) are the same chain; a stage is NS(bn_w, bn_b, rm, rv, w (bf16 shadow [Cout, Cin]), b (fp32 or None)).
def _pad8(n):
    return (n + 7) // 8 * 8


def mlp_head_fwd(x, stages, training, save=True):
    """x fp32 [B, C0] -> fp32 [B, Cout_last] (a view of a buffer whose row stride is padded to 8 columns)."""
    B = x.shape[0]
    ctxs = []
    for st in stages:
        a, bst = ops.bn_forward(x, st.bn_w, st.bn_b, st.rm, st.rv, training, True)
        cout = st.w.shape[0]
        ybuf = _empty((B, _pad8(cout)), F32, x)
        y = ybuf[:, :cout]
        ops.gemm(a, st.w, y, bias=st.b)
        if save:
            ctxs.append(NS(x=x, a=a, st=bst))
        x = y
    return x, ctxs


def mlp_head_bwd(dy, ctxs, stages, grads):
    """dy fp32 [B, Cout_last] -> fp32 [B, C0]; parameter gradients accumulate into grads[i] = NS(bn_w, bn_b, w, b)."""
    B = dy.shape[0]
    g = None
    for i in range(len(stages) - 1, -1, -1):
        st, G, c = stages[i], grads[i], ctxs[i]
        cout, cin = st.w.shape
        if g is None:     # fp32 upstream of the last Linear -> bf16 GEMM operand in a buffer with 16-byte aligned rows
            gp = ops.zeros_(_empty((B, _pad8(cout)), F32, dy)) if cout % 8 else _empty((B, cout), F32, dy)
            ops.copy2d(dy, gp[:, :cout])
            gfull = ops.dropout_grad(gp, 0.0, None, 0)
            g = gfull[:, :cout]
        if G.b is not None:
            if cout % 8 == 0:
                ops.colsum(g, sum32=G.b)
            else:       # padded operand buffer: column sums of the padded width, then the valid part into the bias gradient
                tmp = ops.zeros_(_empty((_pad8(cout),), F32, dy))
                ops.colsum(gfull, sum32=tmp)
                ops.add_scale(tmp[:cout], G.b, 1.0, out=G.b)
        _wgrad(g, c.a, G.w)
        da = _empty((B, cin), BF16, dy)
        _dgrad(g, st.w, da)
        g = ops.bn_backward(da, c.x, c.st, True, G.bn_w, G.bn_b, out_dtype=F32 if i == 0 else BF16)
    return g


def pool_head_fwd(x, W, bn, B, L, D, training, save=True):
    """x fp32 [B*L, D] -> (feats fp32 [B, D], backbone fp32 [B, 2D])."""
    pooled, am = ops.token_pool_fwd(x, B, L, D)
    stages = [NS(bn_w=W.bn1_w, bn_b=W.bn1_b, rm=bn.rm1, rv=bn.rv1, w=W.wa, b=None),
              NS(bn_w=W.bn2_w, bn_b=W.bn2_b, rm=bn.rm2, rv=bn.rv2, w=W.wb, b=None)]
    feats, cs = mlp_head_fwd(pooled, stages, training, save)
    ctx = NS(pooled=pooled, am=am, a1=cs[0].a, a2=cs[1].a, cs=cs, stages=stages) if save else None
    return feats, pooled, ctx


def pool_head_bwd(dfeats, dbackbone, c, W, G, B, L, D):
    if dfeats is not None:
        grads = [NS(bn_w=G.bn1_w, bn_b=G.bn1_b, w=G.wa, b=None), NS(bn_w=G.bn2_w, bn_b=G.bn2_b, w=G.wb, b=None)]
        dpooled = mlp_head_bwd(dfeats, c.cs, c.stages, grads)
        if dbackbone is not None:
            dpooled = ops.add_scale(dpooled, dbackbone.contiguous(), 1.0)
    else:
        dpooled = dbackbone.contiguous()
    return ops.token_pool_bwd(dpooled, c.am, B, L, D)


def pool_cls_head_fwd(x, stages, B, L, D, training, save=True):
    """Fine-tune classifier (partseg.py:597-604): x fp32 [B*L, D] -> logits fp32 [B, classes]."""
    pooled, am = ops.token_pool_fwd(x, B, L, D)
    logits, cs = mlp_head_fwd(pooled, stages, training, save)
    return logits, (NS(am=am, cs=cs, a=[c.a for c in cs]) if save else None)


def pool_cls_head_bwd(dlogits, c, stages, grads, B, L, D):
    return ops.token_pool_bwd(mlp_head_bwd(dlogits, c.cs, stages, grads), c.am, B, L, D)
