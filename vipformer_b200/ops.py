"""Low-level Python bindings of the C-ABI kernels (include/vpf.h).  No autograd here."""
import ctypes

import torch

from . import _lib

EPI_STORE, EPI_RESIDUAL, EPI_ATOMIC_ADD = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
AUX_NONE, AUX_GELU_GRAD, AUX_RELU_MASK = 0, 1, 2

_i = ctypes.c_int
_f = ctypes.c_float
_vp = ctypes.c_void_p


class GemmEpilogue(ctypes.Structure):
    _fields_ = [("mode", _i), ("out_f32", _i), ("ldc", _i), ("act", _i), ("aux_mode", _i), ("ld_aux", _i),
                ("rg_shift", _i), ("rg_ld", _i), ("op_id", ctypes.c_uint), ("alpha", _f), ("drop_p", _f),
                ("out", _vp), ("out2", _vp), ("out_bf16", _vp), ("bias", _vp), ("rg_bias", _vp), ("aux", _vp),
                ("resid", _vp), ("seed_ptr", _vp), ("gm_S", _i), ("gm_ld", _i), ("gm_out_f32", _vp), ("gm_out_bf16", _vp),
                ("gm_argmax", _vp), ("gm_cols", _i), ("row_bias", _vp)]


def _dp(t):
    return None if t is None else t.data_ptr()


def gemm(a, b, out, *, a_mn=False, b_mn=False, M=None, N=None, K=None, lda=None, ldb=None, mode=EPI_STORE,
         bias=None, rg_bias=None, rg_shift=0, act=ACT_NONE, aux=None, aux_mode=AUX_NONE, out2=None, resid=None,
         out_bf16=None, alpha=1.0, drop_p=0.0, seed=None, op_id=0, splits=None, ldc=None, gm_S=0, gm_f32=None,
         gm_bf16=None, gm_argmax=None, gm_cols=False, row_bias=None):
    """out (op)= epilogue(alpha * A @ B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn); bf16."""
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and (out is None or out.stride(-1) == 1)
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    ldc = (out.stride(0) if out is not None else 0) if ldc is None else ldc
    e = GemmEpilogue()
    e.mode, e.out_f32, e.ldc, e.act = mode, int(out is not None and out.dtype == torch.float32), ldc, act
    if mode != EPI_STORE:
        assert out.dtype == torch.float32
    else:
        assert out is None or out.dtype in (torch.float32, torch.bfloat16)
    if gm_S:
        gm_any = gm_f32 if gm_f32 is not None else gm_bf16
        e.gm_S, e.gm_ld = gm_S, gm_any.stride(0)
        e.gm_out_f32, e.gm_out_bf16, e.gm_argmax = _dp(gm_f32), _dp(gm_bf16), _dp(gm_argmax)
        e.gm_cols, e.row_bias = int(gm_cols), _dp(row_bias)
    e.aux_mode, e.ld_aux = aux_mode, (aux.stride(0) if aux is not None else 0)
    e.rg_shift, e.rg_ld = rg_shift, (rg_bias.stride(0) if rg_bias is not None else 0)
    e.op_id, e.alpha, e.drop_p = op_id, alpha, drop_p
    e.out, e.out2, e.out_bf16 = _dp(out), _dp(out2), _dp(out_bf16)
    e.bias, e.rg_bias, e.aux, e.resid, e.seed_ptr = _dp(bias), _dp(rg_bias), _dp(aux), _dp(resid), _dp(seed)
    if splits is None:
        splits = 0 if mode == EPI_ATOMIC_ADD else 1
    _lib.call("vpf_gemm_bf16", _vp(a.data_ptr()), _i(int(a_mn)), _i(lda), _vp(b.data_ptr()), _i(int(b_mn)), _i(ldb),
              _i(M), _i(N), _i(K), _i(splits), ctypes.byref(e), _lib.stream_ptr())
    return out


# ----------------------------------------------------------------------------- helpers
_ll = ctypes.c_longlong
_u = ctypes.c_uint
BF16, F32 = torch.bfloat16, torch.float32


def _p(t):
    return _vp(0 if t is None else t.data_ptr())


def _isbf(t):
    return _i(int(t.dtype == BF16))


def _s():
    return _lib.stream_ptr()


def zeros_(t):
    _lib.call("vpf_fill_zero", _p(t), _ll(t.numel() * t.element_size()), _s())
    return t


def cast_bf16(x, out=None):
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=BF16, device=x.device) if out is None else out
    _lib.call("vpf_cast_bf16", _p(x), _p(out), _ll(x.numel()), _s())
    return out


def gelu_fwd(z):
    h = torch.empty_like(z)
    _lib.call("vpf_gelu_fwd", _p(z), _p(h), _ll(z.numel()), _s())
    return h


def gelu_bwd(dh, z, colsum=None):
    R, C = z.shape
    dz = torch.empty_like(z)
    _lib.call("vpf_gelu_bwd", _p(dh), _p(z), _p(dz), _p(colsum), _ll(R), _i(C), _s())
    return dz


def add_scale(a, b, alpha, out=None):
    out = torch.empty_like(a) if out is None else out
    _lib.call("vpf_add_scale", _p(a), _p(b), _p(out), _f(alpha), _ll(a.numel()), _s())
    return out


def bn_fold(W, b, scale, shift):
    """-> (bf16 [N, K] = scale[:, None] * W, fp32 [N] = scale * b + shift): eval-mode BatchNorm folded into its Linear."""
    N, K = W.shape
    Wo = torch.empty((N, K), dtype=BF16, device=W.device)
    bo = torch.empty(N, dtype=F32, device=W.device)
    _lib.call("vpf_bn_fold", _p(W.contiguous()), _p(b), _p(scale), _p(shift), _p(Wo), _p(bo), _i(N), _i(K), _s())
    return Wo, bo


def copy2d(src, dst, alpha=1.0):
    """dst[:, :] = alpha * src over equal [rows, cols] windows of row-strided fp32 tensors."""
    rows, cols = src.shape
    assert dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1 and src.dtype == F32 and dst.dtype == F32
    _lib.call("vpf_copy2d", _p(src), _i(src.stride(0)), _p(dst), _i(dst.stride(0)), _ll(rows), _i(cols), _f(alpha), _s())
    return dst


def ce_label_smoothing(logits, labels, eps):
    """-> (loss fp32 [1], dlogits fp32 [n, C] = d loss / d logits).  logits fp32 [n, C] (row-strided view allowed)."""
    n, C = logits.shape
    assert logits.dtype == F32 and logits.stride(1) == 1 and labels.dtype == torch.int64 and labels.numel() == n
    loss = zeros_(torch.empty(1, dtype=F32, device=logits.device))
    d = torch.empty((n, C), dtype=F32, device=logits.device)
    _lib.call("vpf_ce_ls", _p(logits), _i(logits.stride(0)), _p(labels.contiguous()), _i(n), _i(C), _f(eps), _p(loss), _p(d), _i(C), _s())
    return loss, d


# ----------------------------------------------------------------------------- attention
def attention_fwd(q, k, v, B, H, Lq, Lk, scale, drop_p=0.0, seed=None, op_id=0):
    """q: [B*Lq, >=H*64] view (row stride = q.stride(0)); k, v: views of one buffer with the same row stride."""
    assert k.stride(0) == v.stride(0)
    o = torch.empty((B * Lq, H * 64), dtype=BF16, device=q.device)
    lse = torch.empty((B * H, Lq), dtype=F32, device=q.device)
    _lib.call("vpf_attention_fwd", _p(q), _i(q.stride(0)), _p(k), _p(v), _i(k.stride(0)), _p(o), _i(o.stride(0)),
              _p(lse), _i(B), _i(H), _i(Lq), _i(Lk), _i(64), _f(scale), _f(drop_p), _p(seed), _u(op_id), _s())
    return o, lse


def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Lq, Lk, scale, drop_p=0.0, seed=None, op_id=0):
    assert k.stride(0) == v.stride(0) and dk.stride(0) == dv.stride(0)
    delta = torch.empty(_ws("vpf_attention_bwd_workspace_bytes", B, H, Lq) // 4, dtype=F32, device=q.device)
    _lib.call("vpf_attention_bwd", _p(q), _i(q.stride(0)), _p(k), _p(v), _i(k.stride(0)), _p(o), _i(o.stride(0)),
              _p(do), _i(do.stride(0)), _p(lse), _p(delta), _p(dq), _i(dq.stride(0)), _p(dk), _p(dv),
              _i(dk.stride(0)), _i(B), _i(H), _i(Lq), _i(Lk), _i(64), _f(scale), _f(drop_p), _p(seed), _u(op_id), _s())


# ----------------------------------------------------------------------------- norms
def layernorm_fwd(x, gamma, beta, *, add=None, want_xsum=False, relu=False, eps=1e-5):
    T, D = x.shape
    y = torch.empty((T, D), dtype=BF16, device=x.device)
    mean = torch.empty(T, dtype=F32, device=x.device)
    rstd = torch.empty(T, dtype=F32, device=x.device)
    xsum = torch.empty((T, D), dtype=F32, device=x.device) if want_xsum else None
    add_rows = 0 if add is None else add.numel() // D
    _lib.call("vpf_layernorm_fwd", _p(x), _isbf(x), _p(add), _i(add_rows), _p(xsum), _p(gamma), _p(beta), _p(y),
              _p(mean), _p(rstd), _i(T), _i(D), _f(eps), _i(int(relu)), _s())
    return y, mean, rstd, xsum


def layernorm_bwd(dy, x, mean, rstd, gamma, *, y_relu=None, dres=None, dgamma=None, dbeta=None, dpos=None,
                  out_dtype=F32):
    T, D = x.shape
    dx = torch.empty((T, D), dtype=out_dtype, device=x.device)
    pos_rows = 0 if dpos is None else dpos.numel() // D
    _lib.call("vpf_layernorm_bwd", _p(dy), _isbf(dy), _p(x), _isbf(x), _p(y_relu), _p(mean), _p(rstd), _p(gamma),
              _p(dres), _p(dx), _isbf(dx), _p(dgamma), _p(dbeta), _p(dpos), _i(pos_rows), _i(T), _i(D), _s())
    return dx


def layernorm_bwd_emit(dy, x, mean, rstd, gamma, emit, *, dres=None, dgamma=None, dbeta=None, dpos=None):
    """layernorm_bwd (bf16 dy, fp32 x -> fp32 dx) that also returns g = bf16(dropout_mask(dx)) for the next block down the
    chain.  emit = (drop_p, seed, op_id, colsum or None).  -> (dx fp32, g bf16)."""
    T, D = x.shape
    p_drop, seed, op_id, colsum = emit
    dx = torch.empty((T, D), dtype=F32, device=x.device)
    g = torch.empty((T, D), dtype=BF16, device=x.device)
    pos_rows = 0 if dpos is None else dpos.numel() // D
    _lib.call("vpf_layernorm_bwd_emit", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dgamma), _p(dbeta),
              _p(dpos), _i(pos_rows), _i(T), _i(D), _p(g), _p(colsum), _f(p_drop), _p(seed), _u(op_id), _s())
    return dx, g


def can_emit(dy, x, D):
    return dy.dtype == BF16 and x.dtype == F32 and D % 128 == 0 and x.numel() < (1 << 32)


def dropout_grad(g, p, seed, op_id, colsum=None):
    T, N = g.shape
    out = torch.empty((T, N), dtype=BF16, device=g.device)
    _lib.call("vpf_dropout_grad", _p(g), _p(out), _p(colsum), _f(p), _p(seed), _u(op_id), _i(T), _i(N), _s())
    return out


def colsum(x, *, sum64=None, sumsq64=None, sum32=None):
    R, C = x.shape
    _lib.call("vpf_colsum", _p(x), _isbf(x), _p(sum64), _p(sumsq64), _p(sum32), _ll(R), _i(C), _s())


class BNState:
    """Folded affine + saved statistics of one BatchNorm1d call."""
    __slots__ = ("scale", "shift", "mean", "rstd", "R", "training")


def bn_stats_finalize(stats, R, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    C = gamma.numel()
    st = BNState()
    buf = torch.empty((4, C), dtype=F32, device=gamma.device)
    st.scale, st.shift, st.mean, st.rstd, st.R, st.training = buf[0], buf[1], buf[2], buf[3], R, bool(training)
    _lib.call("vpf_bn_finalize", _p(stats), _ll(R), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
              _f(momentum), _f(eps), _i(int(training)), _p(st.scale), _p(st.shift), _p(st.mean), _p(st.rstd), _i(C), _s())
    return st


def bn_forward(x, gamma, beta, running_mean, running_var, training, relu, out_dtype=BF16, momentum=0.1, eps=1e-5):
    R, C = x.shape
    stats = None
    if training:
        stats = zeros_(torch.empty(2 * C, dtype=torch.float64, device=x.device))
        colsum(x, sum64=stats[:C], sumsq64=stats[C:])
    st = bn_stats_finalize(stats, R, gamma, beta, running_mean, running_var, training, momentum, eps)
    y = torch.empty((R, C), dtype=out_dtype, device=x.device)
    _lib.call("vpf_bn_apply", _p(x), _isbf(x), _p(st.scale), _p(st.shift), _p(y), _isbf(y), _i(int(relu)), _ll(R), _i(C), _s())
    return y, st


def bn_backward(dy, x, st, relu, dgamma, dbeta, out_dtype=BF16):
    if not st.training:
        # vpf_bn_bwd applies the batch-statistics gradient (mean / projection terms); a forward that normalised with the
        # running statistics (module.eval()) has a different Jacobian.  Every reference script trains with BN in train mode.
        raise NotImplementedError("backward through an eval-mode BatchNorm1d (running statistics) is not built; "
                                  "call .train() on the module before a forward you backpropagate through")
    R, C = x.shape
    red = torch.empty(_ws("vpf_bn_bwd_workspace_bytes", C) // 8, dtype=torch.float64, device=x.device)
    dx = torch.empty((R, C), dtype=out_dtype, device=x.device)
    _lib.call("vpf_bn_bwd", _p(dy), _isbf(dy), _p(x), _isbf(x), _p(st.scale), _p(st.shift), _p(st.mean), _p(st.rstd),
              _i(int(relu)), _p(red), _p(dx), _isbf(dx), _p(dgamma), _p(dbeta), _ll(R), _i(C), _s())
    return dx


def bn_backward_gsum(dy, x, st, relu, dgamma, dbeta, S, want_bf16=True, want_f32=False):
    """bn_backward of an all-bf16 [R, 256] tensor + the sum of dx over groups of S consecutive rows, one pass.
    -> (dx bf16 [R, 256], gsum bf16 [R/S, 256] or None, gsum fp32 or None)."""
    if not st.training:
        raise NotImplementedError("backward through an eval-mode BatchNorm1d (running statistics) is not built")
    R, C = x.shape
    red = torch.empty(_ws("vpf_bn_bwd_workspace_bytes", C) // 8, dtype=torch.float64, device=x.device)
    dx = torch.empty((R, C), dtype=BF16, device=x.device)
    gb = torch.empty((R // S, C), dtype=BF16, device=x.device) if want_bf16 else None
    gf = torch.empty((R // S, C), dtype=F32, device=x.device) if want_f32 else None
    _lib.call("vpf_bn_bwd_gsum", _p(dy), _p(x), _p(st.scale), _p(st.shift), _p(st.mean), _p(st.rstd), _i(int(relu)), _p(red),
              _p(dx), _p(dgamma), _p(dbeta), _ll(R), _i(C), _i(S), _p(gb), _p(gf), _s())
    return dx, gb, gf


# ----------------------------------------------------------------------------- pooling / thin ops
def group_max_fwd(x, G, S, C, want_bf16=True, want_f32=False):
    ob = torch.empty((G, C), dtype=BF16, device=x.device) if want_bf16 else None
    of = torch.empty((G, C), dtype=F32, device=x.device) if want_f32 else None
    am = torch.empty((G, C), dtype=torch.uint8, device=x.device)
    _lib.call("vpf_group_max_fwd", _p(x), _p(ob), _p(of), _p(am), _i(G), _i(S), _i(C), _s())
    return ob, of, am


def group_max_bwd(dout, argmax, G, S, C, dx=None):
    accumulate = dx is not None
    if dx is None:
        dx = torch.empty((G * S, C), dtype=BF16, device=dout.device)
    _lib.call("vpf_group_max_bwd", _p(dout), _isbf(dout), _p(argmax), _p(dx), _i(int(accumulate)), _i(G), _i(S), _i(C), _s())
    return dx


def group_sum(x, G, S, C, want_bf16=True, want_f32=False):
    ob = torch.empty((G, C), dtype=BF16, device=x.device) if want_bf16 else None
    of = torch.empty((G, C), dtype=F32, device=x.device) if want_f32 else None
    _lib.call("vpf_group_sum", _p(x), _p(ob), _p(of), _i(G), _i(S), _i(C), _s())
    return ob, of


def droppath_scales(seed, op_id, p, B, like):
    s = torch.empty(B, dtype=F32, device=like.device)
    _lib.call("vpf_droppath_scales", _p(seed), _u(op_id), _f(p), _i(B), _p(s), _s())
    return s


def row_scale(x, scales, L, out=None):
    """out[r, :] = x[r, :] * scales[r // L]   (fp32 [T, D]; out may be x itself)."""
    T, D = x.shape
    out = torch.empty_like(x) if out is None else out
    assert x.dtype == F32 and x.is_contiguous() and out.is_contiguous()
    _lib.call("vpf_row_scale", _p(x), _p(scales), _i(L), _p(out), _ll(T), _i(D), _s())
    return out


def vecmat_bf16(v, W, out):
    """out[n] += sum_k v[k] * W[k, n];  v fp32 [K], W bf16 [K, N] window (row stride W.stride(0)), out fp32 [N]."""
    K, N = W.shape
    assert v.dtype == F32 and W.dtype == BF16 and out.dtype == F32 and v.numel() == K and out.numel() == N and W.stride(1) == 1
    _lib.call("vpf_vecmat_bf16", _p(v), _p(W), _i(W.stride(0)), _i(K), _i(N), _p(out), _s())
    return out


def token_pool_fwd(x, B, L, D):
    out = torch.empty((B, 2 * D), dtype=F32, device=x.device)
    am = torch.empty((B, D), dtype=torch.int32, device=x.device)
    _lib.call("vpf_token_pool_fwd", _p(x), _p(out), _p(am), _i(B), _i(L), _i(D), _s())
    return out, am


def token_pool_bwd(dout, am, B, L, D):
    dx = torch.empty((B * L, D), dtype=F32, device=dout.device)
    _lib.call("vpf_token_pool_bwd", _p(dout), _p(am), _p(dx), _i(B), _i(L), _i(D), _s())
    return dx


def linear3_fwd(p, ldp, w, b, R, *, scale=None, shift=None, want_pre=False, act=None):
    Co = w.shape[0]
    pre = torch.empty((R, Co), dtype=BF16, device=p.device) if want_pre else None
    ao = torch.empty((R, Co), dtype=BF16, device=p.device) if act is not None else None
    _lib.call("vpf_linear3_fwd", _p(p), _i(ldp), _p(w), _p(b), _p(scale), _p(shift), _p(pre), _p(ao),
              _i(act if act is not None else 0), _ll(R), _i(Co), _s())
    return pre, ao


def linear3_stats(p, ldp, w, b, R):
    Co = w.shape[0]
    stats = zeros_(torch.empty(2 * Co, dtype=torch.float64, device=p.device))
    _lib.call("vpf_linear3_stats", _p(p), _i(ldp), _p(w), _p(b), _p(stats), _ll(R), _i(Co), _s())
    return stats


def linear3_bwd(dy, p, ldp, dW, db, R):
    _lib.call("vpf_linear3_bwd", _p(dy), _isbf(dy), _p(p), _i(ldp), _p(dW), _p(db), _ll(R), _i(dW.shape[0]), _s())


def linear3_bn_bwd(dh, p, ldp, w, b, st, dW, db, dgamma, dbeta, R):
    Co = w.shape[0]
    red = torch.empty(_ws("vpf_linear3_bn_bwd_workspace_bytes", Co) // 8, dtype=torch.float64, device=p.device)
    _lib.call("vpf_linear3_bn_bwd", _p(dh), _p(p), _i(ldp), _p(w), _p(b), _p(st.scale), _p(st.shift), _p(st.mean),
              _p(st.rstd), _p(red), _p(dW), _p(db), _p(dgamma), _p(dbeta), _ll(R), _i(Co), _s())


def patchify(img, P, nchw=False):
    if nchw:
        B, Ci, H, W = img.shape
    else:
        B, H, W, Ci = img.shape
    out = torch.empty((B * (H // P) * (W // P), P * P * Ci), dtype=BF16, device=img.device)
    _lib.call("vpf_patchify", _p(img), _p(out), _i(B), _i(H), _i(W), _i(Ci), _i(P), _i(int(nchw)), _s())
    return out


# ----------------------------------------------------------------------------- loss / optimiser
def l2norm_rows(x, z=None, norm=None):
    n, D = x.shape
    z = torch.empty_like(x) if z is None else z
    norm = torch.empty(n, dtype=F32, device=x.device) if norm is None else norm
    _lib.call("vpf_l2norm_rows", _p(x), _p(z), _p(norm), _i(n), _i(D), _s())
    return z, norm


def ntxent_fwd(zr, zc, b_local, col_offset, half, temperature, loss_out, n_c=None, colmap=None, lse_out=None):
    """-> (lse [n_r], logits scratch S [n_r, n_c] to hand to ntxent_bwd).  colmap = (blk, ld, base) maps logits column j
    to row (j // blk) * ld + base + j % blk of zc (default: zc is a plain [n_c, D] buffer)."""
    n_r, D = zr.shape
    n_c = zc.shape[0] if n_c is None else n_c
    blk, ld, base = (n_c, n_c, 0) if colmap is None else colmap
    lse = torch.empty(n_r, dtype=F32, device=zr.device) if lse_out is None else lse_out
    S = torch.empty(_ws("vpf_ntxent_logits_workspace_bytes", n_r, n_c) // 4, dtype=F32, device=zr.device).view(n_r, n_c)
    _lib.call("vpf_ntxent_fwd", _p(zr), _i(n_r), _p(zc), _i(n_c), _i(D), _i(b_local), _i(col_offset), _i(half),
              _i(blk), _i(ld), _i(base), _f(temperature), _p(S), _p(lse), _p(loss_out), _s())
    return lse, S


def ntxent_bwd(zr, norm, zc, lse_all, S, b_local, col_offset, half, temperature, gscale, upstream=None, n_c=None,
               colmap=None):
    n_r, D = zr.shape
    n_c = zc.shape[0] if n_c is None else n_c
    blk, ld, base = (n_c, n_c, 0) if colmap is None else colmap
    dx = torch.empty((n_r, D), dtype=F32, device=zr.device)
    G = torch.empty(_ws("vpf_ntxent_grad_workspace_bytes", n_r, D) // 4, dtype=F32, device=zr.device)
    _lib.call("vpf_ntxent_bwd", _p(zr), _p(norm), _i(n_r), _p(zc), _p(lse_all), _i(n_c), _i(D), _i(b_local),
              _i(col_offset), _i(half), _i(blk), _i(ld), _i(base), _f(temperature), _f(gscale), _p(upstream), _p(S), _p(G),
              _p(dx), _s())
    return dx


def ntxent_pack_fwd(z, nseg, zc, b_local, col_offset, half, temperature, n_c, colmap, base_stride):
    """nseg loss terms in one launch per stage: z [nseg * n_r, D] -> (losses [nseg], lse [nseg * n_r], S [nseg, n_r, n_c])."""
    n, D = z.shape
    n_r = n // nseg
    blk, ld, base = colmap
    losses = zeros_(torch.empty(nseg, dtype=F32, device=z.device))
    lse = torch.empty(n, dtype=F32, device=z.device)
    S = torch.empty(nseg * _ws("vpf_ntxent_logits_workspace_bytes", n_r, n_c) // 4, dtype=F32, device=z.device).view(nseg, n_r, n_c)
    _lib.call("vpf_ntxent_pack_fwd", _p(z), _i(nseg), _i(n_r), _p(zc), _i(n_c), _i(D), _i(b_local), _i(col_offset), _i(half),
              _i(blk), _i(ld), _i(base), _i(base_stride), _f(temperature), _p(S), _p(lse), _p(losses), _s())
    return losses, lse, S


def ntxent_pack_bwd(z, norm, nseg, zc, lse_all, S, b_local, col_offset, half, temperature, gscales, upstream, n_c, colmap,
                    base_stride):
    """-> dx [nseg * n_r, D]; gscales: one host float per term."""
    n, D = z.shape
    n_r = n // nseg
    blk, ld, base = colmap
    dx = torch.empty((n, D), dtype=F32, device=z.device)
    G = torch.empty(nseg * _ws("vpf_ntxent_grad_workspace_bytes", n_r, D) // 4, dtype=F32, device=z.device)
    gs = (ctypes.c_float * nseg)(*[float(g) for g in gscales])
    _lib.call("vpf_ntxent_pack_bwd", _p(z), _p(norm), _i(nseg), _i(n_r), _p(zc), _p(lse_all), _i(n_c), _i(D), _i(b_local),
              _i(col_offset), _i(half), _i(blk), _i(ld), _i(base), _i(base_stride), _f(temperature), gs, _p(upstream), _p(S),
              _p(G), _p(dx), _s())
    return dx


def adamw(p, g, m, v, shadow, lr_ptr, step_ptr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, grad_scale=1.0):
    _lib.call("vpf_adamw", _p(p), _p(g), _p(m), _p(v), _p(shadow), _ll(p.numel()), _p(lr_ptr), _f(beta1), _f(beta2),
              _f(eps), _f(weight_decay), _p(step_ptr), _f(grad_scale), _s())


def draw_indices(state, op_id, N, out):
    """out int64 [n] <- uniform indices in [0, N) from the device step state (engine: FPS start points, utils.py:71)."""
    _lib.call("vpf_draw_indices", _p(state), _u(op_id), _i(out.numel()), _i(N), _p(out), _s())
    return out


def _ws(name, *dims):
    return int(_lib.size_query(name, *[_i(d) for d in dims]))


def step_advance(state):
    _lib.call("vpf_step_advance", _p(state), _s())


# ----------------------------------------------------------------------------- part segmentation helpers
def three_nn(pts, centers):
    """pts fp32 [B,N,3], centers fp32 [B,S,3] -> (idx int32 [B,N,3], w fp32 [B,N,3])  (utils.py:223-229)."""
    B, N, _ = pts.shape
    S = centers.shape[1]
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=pts.device)
    w = torch.empty((B, N, 3), dtype=F32, device=pts.device)
    _lib.call("vpf_three_nn", _p(pts), _p(centers), _i(B), _i(N), _i(S), _p(idx), _p(w), _s())
    return idx, w


def interp3_fwd(feats, idx, w, pts, ldo):
    """feats bf16 [B*S, C] -> bf16 [B*N, ldo]: columns [0,C) interpolated, [C,C+3) xyz, rest 0."""
    B, N, _ = pts.shape
    C = feats.shape[1]
    S = feats.shape[0] // B
    out = torch.empty((B * N, ldo), dtype=BF16, device=feats.device)
    _lib.call("vpf_interp3_fwd", _p(feats), _i(feats.stride(0)), _p(idx), _p(w), _p(pts), _p(out), _i(ldo), _i(B), _i(N), _i(S),
              _i(C), _s())
    return out


def interp3_bwd(dout, idx, w, dfeats, B, N, S, C):
    _lib.call("vpf_interp3_bwd", _p(dout), _i(dout.stride(0)), _p(idx), _p(w), _p(dfeats), _i(dfeats.stride(0)), _i(B), _i(N),
              _i(S), _i(C), _s())


def token_pool_bf16_fwd(x, B, L, C):
    out = torch.empty((B, 2 * C), dtype=F32, device=x.device)
    am = torch.empty((B, C), dtype=torch.int32, device=x.device)
    _lib.call("vpf_token_pool_bf16_fwd", _p(x), _i(x.stride(0)), _p(out), _p(am), _i(B), _i(L), _i(C), _s())
    return out, am


def token_pool_accum_bwd(dout, am, dx, B, L, C):
    _lib.call("vpf_token_pool_accum_bwd", _p(dout), _p(am), _p(dx), _i(dx.stride(0)), _i(B), _i(L), _i(C), _s())


def leaky_relu_fwd(x, slope):
    y = torch.empty_like(x)
    _lib.call("vpf_leaky_relu_fwd", _p(x), _p(y), _f(slope), _ll(x.numel()), _s())
    return y


def leaky_relu_bwd(dy, x, slope):
    dx = torch.empty_like(x)
    _lib.call("vpf_leaky_relu_bwd", _p(dy), _p(x), _p(dx), _f(slope), _ll(x.numel()), _s())
    return dx


def permute_w(W, C, ldo):
    Co = W.shape[0]
    out = torch.empty((Co, ldo), dtype=BF16, device=W.device)
    _lib.call("vpf_permute_w", _p(W.contiguous()), _p(out), _i(Co), _i(C), _i(ldo), _s())
    return out


def unpermute_dw(dWp, dW, C):
    _lib.call("vpf_unpermute_dw", _p(dWp), _p(dW), _i(dW.shape[0]), _i(C), _i(dWp.stride(0)), _s())


def copy2d_bf16(src, dst):
    rows, cols = src.shape
    assert dst.shape == src.shape and src.dtype == BF16 and dst.dtype == BF16
    _lib.call("vpf_copy2d_bf16", _p(src), _i(src.stride(0)), _p(dst), _i(dst.stride(0)), _ll(rows), _i(cols), _s())
    return dst
