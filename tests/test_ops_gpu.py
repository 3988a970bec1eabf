"""GPU: every non-GEMM kernel against a plain PyTorch fp32 reference of the same op.
Tolerances: fp32-storage outputs rel 1e-4; bf16-storage outputs one bf16 rounding (2^-8 rel of scale);
attention (bf16 tensor-core inputs, fp32 accumulate) rel-Frobenius 1e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def rnd(shape, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(dtype)


def relfro(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def close(a, b, bf16=False, tol=None):
    a, b = a.float(), b.float()
    scale = b.abs().max().item() + 1e-6
    t = tol if tol is not None else (1 / 128 if bf16 else 1e-4)
    err = (a - b).abs().max().item()
    assert err <= t * scale, f"max err {err} vs tol {t * scale}"


# --------------------------------------------------------------------------- attention
def attn_ref(q, k, v, B, H, Lq, Lk, scale):
    qh = q.float().view(B, Lq, H, 64).transpose(1, 2)
    kh = k.float().view(B, Lk, H, 64).transpose(1, 2)
    vh = v.float().view(B, Lk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    p = s.softmax(-1)
    o = (p @ vh).transpose(1, 2).reshape(B * Lq, H * 64)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,H,Lq,Lk", [(3, 4, 128, 128), (2, 6, 144, 144), (2, 4, 96, 96), (2, 4, 128, 2048),
                                       (1, 4, 128, 1000), (2, 2, 40, 75),
                                       # more (sample, head) items than persistent CTAs: several items per CTA, with one
                                       # and with two query blocks, ragged key tails
                                       (50, 4, 144, 144), (80, 4, 128, 300), (90, 4, 200, 130)])
def test_attention_fwd_bwd(B, H, Lq, Lk):
    from vipformer_b200 import ops

    D = H * 64
    scale = 64 ** -0.5
    q = rnd((B * Lq, D), 1, BF16)
    kv = rnd((B * Lk, 2 * D), 2, BF16)
    k, v = kv[:, :D], kv[:, D:]
    o, lse = ops.attention_fwd(q, k, v, B, H, Lq, Lk, scale)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    oref, lref = attn_ref(qf, kf, vf, B, H, Lq, Lk, scale)
    assert relfro(o, oref) < 1e-2
    close(lse.view(B, H, Lq) * math.log(2.0), lref, tol=2e-3)
    do = rnd((B * Lq, D), 3, BF16)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.attention_bwd(q, k, v, o, do, lse, dq, dkv[:, :D], dkv[:, D:], B, H, Lq, Lk, scale)
    oref.backward(do.float())
    assert relfro(dq, qf.grad) < 2e-2
    assert relfro(dkv[:, :D], kf.grad) < 2e-2
    assert relfro(dkv[:, D:], vf.grad) < 2e-2


def test_attention_dropout_consistency():
    """With dropout the forward is linear in V for a fixed mask, so the backward (which regenerates the mask)
    must satisfy <dO, O(V)> == <dV, V>; and the keep rate must be 1-p."""
    from vipformer_b200 import ops

    B, H, L = 90, 4, 144        # two query blocks, two key blocks, several items per CTA
    D = H * 64
    seed = torch.tensor([77], device="cuda", dtype=torch.int64)
    q = rnd((B * L, D), 1, BF16)
    kv = rnd((B * L, 2 * D), 2, BF16)
    k, v = kv[:, :D], kv[:, D:]
    o1, lse = ops.attention_fwd(q, k, v, B, H, L, L, 0.125, 0.1, seed, 5)
    o2, _ = ops.attention_fwd(q, k, v, B, H, L, L, 0.125, 0.1, seed, 5)
    assert torch.equal(o1, o2)
    o0, _ = ops.attention_fwd(q, k, v, B, H, L, L, 0.125)
    assert 0.01 < relfro(o1, o0) < 1.0
    do = rnd((B * L, D), 3, BF16)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.attention_bwd(q, k, v, o1, do, lse, dq, dkv[:, :D], dkv[:, D:], B, H, L, L, 0.125, 0.1, seed, 5)
    lhs = (do.float() * o1.float()).sum().item()
    rhs = (dkv[:, D:].float() * v.float()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * abs(lhs) + 1e-2


@pytest.mark.parametrize("B,H,Lq,Lk", [(3, 4, 128, 128), (40, 4, 144, 144), (2, 4, 128, 300), (2, 2, 160, 40)])
def test_attention_dropout_masks_equal_the_oracle_restatement(B, H, Lq, Lk):
    """Dropout ON, exactly: the keep-mask the kernels regenerate from (seed, op id, element) is restated in
    oracle/rng.py; softmax -> that mask -> . V in plain PyTorch must reproduce forward AND backward of both kernels
    (tcgen05 path for Lq <= 256, one and two query blocks, ragged key tails; legacy mma.sync path beyond)."""
    from oracle import rng as R
    from vipformer_b200 import ops

    D, scale, p, seed_v, op = H * 64, 64 ** -0.5, 0.1, 0x1234ABCD5678, 40
    seed = torch.tensor([seed_v], device="cuda", dtype=torch.int64)
    q = rnd((B * Lq, D), 1, BF16)
    kv = rnd((B * Lk, 2 * D), 2, BF16)
    k, v = kv[:, :D], kv[:, D:]
    o, lse = ops.attention_fwd(q, k, v, B, H, Lq, Lk, scale, p, seed, op)
    keep = torch.from_numpy(R.attention_keep(seed_v, op, p, B * H, Lq, Lk)).cuda().view(B, H, Lq, Lk)
    assert 0.85 < (keep > 0).float().mean().item() < 0.95
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(B, Lq, H, 64).transpose(1, 2)
    kh = kf.view(B, Lk, H, 64).transpose(1, 2)
    vh = vf.view(B, Lk, H, 64).transpose(1, 2)
    pr = (qh @ kh.transpose(-1, -2) * scale).softmax(-1) * keep
    oref = (pr @ vh).transpose(1, 2).reshape(B * Lq, D)
    assert relfro(o, oref) < 1e-2, relfro(o, oref)
    do = rnd((B * Lq, D), 3, BF16)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.attention_bwd(q, k, v, o, do, lse, dq, dkv[:, :D], dkv[:, D:], B, H, Lq, Lk, scale, p, seed, op)
    oref.backward(do.float())
    assert relfro(dq, qf.grad) < 2e-2 and relfro(dkv[:, :D], kf.grad) < 2e-2 and relfro(dkv[:, D:], vf.grad) < 2e-2, \
        (relfro(dq, qf.grad), relfro(dkv[:, :D], kf.grad), relfro(dkv[:, D:], vf.grad))


# --------------------------------------------------------------------------- layernorm
@pytest.mark.parametrize("D", [64, 256, 384, 512, 768])
def test_layernorm_fwd_bwd(D):
    from vipformer_b200 import ops

    T = 1000
    x = rnd((T, D), 1)
    pos = rnd((250, D), 2)
    gamma, beta = rnd(D, 3) * 0.1 + 1, rnd(D, 4) * 0.1
    y, mean, rstd, xsum = ops.layernorm_fwd(x, gamma, beta, add=pos, want_xsum=True)
    xs = (x.view(4, 250, D) + pos).view(T, D)
    close(xsum, xs)
    xr = xs.clone().requires_grad_(True)
    g2, b2 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), g2, b2, 1e-5)
    close(y, yr, bf16=True)
    dy = rnd((T, D), 5)
    dres = rnd((T, D), 6)
    yr.backward(dy)
    dgamma, dbeta = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dpos = torch.zeros_like(pos)
    dx = ops.layernorm_bwd(dy, xsum, mean, rstd, gamma, dres=dres, dgamma=dgamma, dbeta=dbeta, dpos=dpos)
    close(dx, xr.grad + dres, tol=2e-4)
    close(dgamma, g2.grad, tol=1e-3)
    close(dbeta, b2.grad, tol=1e-3)
    close(dpos, (xr.grad + dres).view(4, 250, D).sum(0), tol=1e-3)


def test_layernorm_relu_bf16_path():
    from vipformer_b200 import ops

    T, D = 4096, 64
    x = rnd((T, D), 1, BF16)
    gamma, beta = rnd(D, 3) * 0.1 + 1, rnd(D, 4) * 0.1
    y, mean, rstd, _ = ops.layernorm_fwd(x, gamma, beta, relu=True)
    xr = x.float().requires_grad_(True)
    yr = F.relu(F.layer_norm(xr, (D,), gamma, beta, 1e-5))
    close(y, yr, bf16=True)
    dy = rnd((T, D), 5, BF16)
    mask = (y.float() > 0)
    (F.layer_norm(xr, (D,), gamma, beta, 1e-5) * mask).backward(dy.float())
    dgamma, dbeta = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, gamma, y_relu=y, dgamma=dgamma, dbeta=dbeta, out_dtype=BF16)
    assert relfro(dx, xr.grad) < 1e-2


# --------------------------------------------------------------------------- batchnorm
@pytest.mark.parametrize("R,C,dt", [(4096, 256, BF16), (512, 512, torch.float32), (100000, 64, BF16)])
def test_batchnorm_train_fwd_bwd(R, C, dt):
    from vipformer_b200 import ops

    x = (rnd((R, C), 1) * 2 + 0.5).to(dt)
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(rnd(C, 2) * 0.2 + 1)
        bn.bias.copy_(rnd(C, 3) * 0.2)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    y, st = ops.bn_forward(x, bn.weight.detach(), bn.bias.detach(), rm, rv, True, True, out_dtype=dt)
    xr = x.float().requires_grad_(True)
    yr = F.relu(bn(xr))
    close(y, yr, bf16=(dt == BF16), tol=None if dt == BF16 else 1e-3)
    close(rm, bn.running_mean, tol=1e-3)
    close(rv, bn.running_var, tol=1e-3)
    dy = rnd((R, C), 4).to(dt)
    yr.backward(dy.float())
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.bn_backward(dy, x, st, True, dgamma, dbeta, out_dtype=dt)
    assert relfro(dx, xr.grad) < (2e-2 if dt == BF16 else 1e-3)
    assert relfro(dgamma, bn.weight.grad) < 2e-2
    assert relfro(dbeta, bn.bias.grad) < 2e-2


def test_batchnorm_eval_uses_running_stats():
    from vipformer_b200 import ops

    x = rnd((300, 128), 1)
    bn = torch.nn.BatchNorm1d(128).cuda().eval()
    with torch.no_grad():
        bn.running_mean.copy_(rnd(128, 2))
        bn.running_var.copy_(rnd(128, 3).abs() + 0.5)
    y, _ = ops.bn_forward(x, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, False, False,
                          out_dtype=torch.float32)
    close(y, bn(x), tol=1e-4)


# --------------------------------------------------------------------------- thin ops
def test_dropout_grad_and_colsum():
    from vipformer_b200 import ops

    g = rnd((2048, 256), 1)
    cs = torch.zeros(256, device="cuda")
    out = ops.dropout_grad(g, 0.0, None, 0, colsum=cs)
    close(out, g, bf16=True)
    close(cs, g.sum(0), tol=1e-3)
    seed = torch.tensor([5], device="cuda", dtype=torch.int64)
    out = ops.dropout_grad(g, 0.5, seed, 9)
    keep = out != 0
    assert abs(keep.float().mean().item() - 0.5) < 0.01
    close(out[keep], 2 * g[keep], bf16=True)
    # same stream as the GEMM residual epilogue (mask regenerated in backward)
    A = torch.eye(256, device="cuda", dtype=BF16)
    o = torch.empty((2048, 256), device="cuda")
    ops.gemm(g.to(BF16), A, o, mode=ops.EPI_RESIDUAL, resid=torch.zeros_like(g), drop_p=0.5, seed=seed, op_id=9)
    assert torch.equal(o != 0, keep)
    x = rnd((5000, 96), 2, BF16)
    s = ops.zeros_(torch.empty(192, dtype=torch.float64, device="cuda"))
    ops.colsum(x, sum64=s[:96], sumsq64=s[96:])
    close(s[:96], x.double().sum(0), tol=1e-5)
    close(s[96:], (x.double() ** 2).sum(0), tol=1e-5)


@pytest.mark.parametrize("T,N,p", [(2048, 256, 0.5), (300, 96, 0.1), (77, 384, 0.5), (130, 72, 0.1)])
def test_residual_dropout_mask_equals_the_oracle_restatement(T, N, p):
    """Byte-granular Residual masks (rng.cuh keep8): dropout_grad and the GEMM residual epilogue draw exactly the mask
    oracle/rng.py restates, scale 256 / (256 - round(256 p)) included."""
    import oracle.rng as R
    from vipformer_b200 import ops

    g = rnd((T, N), 3) + 3.0          # no zeros in the data: out == 0 <=> dropped
    seed = torch.tensor([0x1234567 + T], device="cuda", dtype=torch.int64)
    m = torch.from_numpy(R.residual_keep(int(seed.item()), 11, p, T, N)).cuda()
    out = ops.dropout_grad(g, p, seed, 11)
    assert torch.equal(out != 0, m != 0)
    close(out, g * m, bf16=True)
    A = torch.eye(N, device="cuda", dtype=BF16)
    o = torch.empty((T, N), device="cuda")
    ops.gemm(g.to(BF16), A, o, mode=ops.EPI_RESIDUAL, resid=torch.zeros_like(g), drop_p=p, seed=seed, op_id=11)
    assert torch.equal(o != 0, m != 0)
    close(o, g.to(BF16).float() * m, tol=1e-5)
    assert abs((m != 0).float().mean().item() - (1 - R.threshold8(p) / 256)) < 0.02


@pytest.mark.parametrize("T,D,pos_rows,p", [(4096, 256, 128, 0.5), (1152, 384, 144, 0.1), (1000, 256, 0, 0.0), (77, 128, 0, 0.5)])
def test_layernorm_backward_emit_equals_two_kernels(T, D, pos_rows, p):
    """vpf_layernorm_bwd_emit == vpf_layernorm_bwd followed by vpf_dropout_grad: same dx, same masked bf16 copy (bit for
    bit: same mask stream), same bias column sums, same dgamma / dbeta / dpos."""
    from vipformer_b200 import ops

    x, dres = rnd((T, D), 1), rnd((T, D), 2)
    dy = rnd((T, D), 3, BF16)
    gam = rnd((D,), 4) * 0.2 + 1.0
    _, mean, rstd, _ = ops.layernorm_fwd(x, gam, torch.zeros(D, device="cuda"))
    seed = torch.tensor([99 + T], device="cuda", dtype=torch.int64)

    def z(*shape):
        return torch.zeros(shape, device="cuda")

    dg0, db0, cs0 = z(D), z(D), z(D)
    dp0 = z(pos_rows, D) if pos_rows else None
    dx0 = ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dres, dgamma=dg0, dbeta=db0, dpos=dp0)
    g0 = ops.dropout_grad(dx0, p, seed, 21, colsum=cs0)
    dg1, db1, cs1 = z(D), z(D), z(D)
    dp1 = z(pos_rows, D) if pos_rows else None
    dx1, g1 = ops.layernorm_bwd_emit(dy, x, mean, rstd, gam, (p, seed, 21, cs1), dres=dres, dgamma=dg1, dbeta=db1, dpos=dp1)
    assert torch.equal(dx1, dx0) and torch.equal(g1, g0)
    close(cs1, cs0, tol=1e-5)
    close(dg1, dg0, tol=1e-5)
    close(db1, db0, tol=1e-5)
    if pos_rows:
        close(dp1, dp0, tol=1e-5)


@pytest.mark.parametrize("G,S", [(300, 32), (17, 8), (5, 33)])
def test_bn_backward_fused_group_sum(G, S):
    """vpf_bn_bwd_gsum == vpf_bn_bwd followed by vpf_group_sum (Group2Emb backward), dgamma / dbeta included."""
    from vipformer_b200 import ops

    R, C = G * S, 256
    x, dy = rnd((R, C), 1, BF16), rnd((R, C), 2, BF16)
    w, b = rnd((C,), 3) * 0.2 + 1.0, rnd((C,), 4) * 0.1
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    _, st = ops.bn_forward(x, w, b, rm, rv, True, True)
    dg0, db0 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx0 = ops.bn_backward(dy, x, st, True, dg0, db0)
    sb0, sf0 = ops.group_sum(dx0, G, S, C, want_bf16=True, want_f32=True)
    dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx1, sb1, sf1 = ops.bn_backward_gsum(dy, x, st, True, dg1, db1, S, want_bf16=True, want_f32=True)
    close(dx1, dx0, bf16=True)
    close(dg1, dg0, tol=1e-5)
    close(db1, db0, tol=1e-5)
    # the fused sum adds the fp32 values before their bf16 rounding; the two-kernel path sums the rounded tensor
    close(sf1, dx0.float().view(G, S, C).sum(1), tol=2e-2)
    close(sf1, dx1.float().view(G, S, C).sum(1), tol=2e-2)
    close(sb1, sf1, bf16=True)
    close(sf0, sf1, tol=2e-2)


def test_group_max_and_token_pool():
    from vipformer_b200 import ops

    G, S, C = 300, 32, 128
    x = rnd((G * S, C), 1, BF16)
    ob, of, am = ops.group_max_fwd(x, G, S, C, True, True)
    ref, ri = x.float().view(G, S, C).max(1)
    assert torch.equal(of, ref) and torch.equal(ob.float(), ref)
    assert torch.equal(am.long(), ri)
    dout = rnd((G, C), 2)
    dx = ops.group_max_bwd(dout, am, G, S, C)
    exp = torch.zeros((G, S, C), device="cuda").scatter_(1, ri[:, None], dout[:, None])
    close(dx.view(G, S, C), exp, bf16=True)
    dx2 = ops.group_max_bwd(dout, am, G, S, C, dx=dx.clone())
    close(dx2.view(G, S, C), 2 * exp, bf16=True, tol=1 / 64)
    B, L, D = 7, 144, 384
    t = rnd((B * L, D), 3)
    out, am = ops.token_pool_fwd(t, B, L, D)
    tr = t.view(B, L, D).clone().requires_grad_(True)
    ref = torch.cat([tr.max(1)[0], tr.mean(1)], 1)
    close(out, ref)
    dout = rnd((B, 2 * D), 4)
    ref.backward(dout)
    close(ops.token_pool_bwd(dout, am, B, L, D).view(B, L, D), tr.grad)


def test_linear3_family():
    from vipformer_b200 import ops

    R, Co = 20000, 64
    p = rnd((R, 3), 1)
    w, b = rnd((Co, 3), 2), rnd(Co, 3)
    y = p @ w.t() + b
    pre, act = ops.linear3_fwd(p, 3, w, b, R, want_pre=True, act=ops.ACT_GELU)
    close(pre, y, bf16=True)
    close(act, F.gelu(y), bf16=True)
    st64 = ops.linear3_stats(p, 3, w, b, R)
    close(st64[:Co], y.double().sum(0), tol=1e-4)
    close(st64[Co:], (y.double() ** 2).sum(0), tol=1e-4)
    dy = rnd((R, Co), 4, BF16)
    dW, db = torch.zeros((Co, 3), device="cuda"), torch.zeros(Co, device="cuda")
    ops.linear3_bwd(dy, p, 3, dW, db, R)
    close(dW, dy.float().t() @ p, tol=1e-3)
    close(db, dy.float().sum(0), tol=1e-3)
    # conv(3->64) + BN(train) + ReLU, fused forward and backward, vs autograd
    bn = torch.nn.BatchNorm1d(Co).cuda()
    with torch.no_grad():
        bn.weight.copy_(rnd(Co, 5) * 0.2 + 1)
        bn.bias.copy_(rnd(Co, 6) * 0.2)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    hr = F.relu(bn(p @ wr.t() + br))
    st = ops.bn_stats_finalize(st64, R, bn.weight.detach(), bn.bias.detach(), None, None, True)
    _, h = ops.linear3_fwd(p, 3, w, b, R, scale=st.scale, shift=st.shift, act=ops.ACT_RELU)
    close(h, hr, bf16=True)
    dh = rnd((R, Co), 7, BF16)
    hr.backward(dh.float())
    dW.zero_(); db.zero_()
    dg, dbt = torch.zeros(Co, device="cuda"), torch.zeros(Co, device="cuda")
    ops.linear3_bn_bwd(dh, p, 3, w, b, st, dW, db, dg, dbt, R)
    assert relfro(dW, wr.grad) < 5e-3
    # db is analytically 0 (BN removes the mean): only noise on the scale of dW's rounding is allowed
    assert db.abs().max().item() < 1e-3 * wr.grad.abs().max().item() and br.grad.abs().max().item() < 1e-3 * wr.grad.abs().max().item()
    assert relfro(dg, bn.weight.grad) < 5e-3 and relfro(dbt, bn.bias.grad) < 5e-3


def test_patchify_and_misc():
    from einops import rearrange
    from vipformer_b200 import ops

    img = rnd((3, 144, 144, 3), 1)
    out = ops.patchify(img, 12)
    ref = rearrange(img, "b (h p1) (w p2) c -> (b h w) (p1 p2 c)", p1=12, p2=12)
    assert torch.equal(out.float(), ref.to(BF16).float())
    # NCHW source (what the data loader yields before pretrain.py:179 permutes it): same patches, bit for bit
    out2 = ops.patchify(img.permute(0, 3, 1, 2).contiguous(), 12, nchw=True)
    assert torch.equal(out2, out)
    a, b = rnd(1000, 2), rnd(1000, 3)
    close(ops.add_scale(a, b, 0.5), 0.5 * (a + b))
    close(ops.cast_bf16(a), a, bf16=True)


# --------------------------------------------------------------------------- loss / optimiser
def ntxent_ref(out0, out1, T):
    """lightly 1.1.21 NTXentLoss.forward restated (see oracle/model_ref.py)."""
    z = torch.cat([F.normalize(out0, dim=1), F.normalize(out1, dim=1)], 0)
    n = z.shape[0]
    logits = z @ z.t() / T
    logits = logits[~torch.eye(n, dtype=torch.bool, device=z.device)].view(n, n - 1)
    b = n // 2
    labels = torch.cat([torch.arange(b, device=z.device) + b - 1, torch.arange(b, device=z.device)])
    return F.cross_entropy(logits, labels)


@pytest.mark.parametrize("b,D", [(55, 256), (256, 256), (64, 384)])
def test_ntxent_fwd_bwd(b, D):
    from vipformer_b200 import ops

    x0, x1 = rnd((b, D), 1), rnd((b, D), 2)
    x = torch.cat([x0, x1], 0).contiguous()
    z, norm = ops.l2norm_rows(x)
    loss = torch.zeros(1, device="cuda")
    lse, S = ops.ntxent_fwd(z, z, b, 0, b, 0.1, loss)
    xr0, xr1 = x0.clone().requires_grad_(True), x1.clone().requires_grad_(True)
    lref = ntxent_ref(xr0, xr1, 0.1)
    assert abs(loss.item() - lref.item()) < 1e-4 * max(1.0, abs(lref.item()))
    lref.backward()
    dx = ops.ntxent_bwd(z, norm, z, lse, S, b, 0, b, 0.1, 1.0 / (2 * b))
    close(dx[:b], xr0.grad, tol=1e-3)
    close(dx[b:], xr1.grad, tol=1e-3)


def test_ntxent_global_negatives_matches_single_process():
    """Two 'ranks' each holding half the batch, columns = all-gathered embeddings, must reproduce the loss and the
    gradient of one process on the concatenated batch."""
    from vipformer_b200 import ops

    b, W, D = 32, 2, 256
    x0, x1 = rnd((b * W, D), 1), rnd((b * W, D), 2)
    xr0, xr1 = x0.clone().requires_grad_(True), x1.clone().requires_grad_(True)
    lref = ntxent_ref(xr0, xr1, 0.1)
    lref.backward()
    zc, _ = ops.l2norm_rows(torch.cat([x0, x1], 0).contiguous())
    losses, lses, locals_ = [], [], []
    for r in range(W):
        xl = torch.cat([x0[r * b:(r + 1) * b], x1[r * b:(r + 1) * b]], 0).contiguous()
        z, norm = ops.l2norm_rows(xl)
        loss = torch.zeros(1, device="cuda")
        lse, S = ops.ntxent_fwd(z, zc, b, r * b, b * W, 0.1, loss)
        losses.append(loss)
        lses.append(lse)
        locals_.append((z, norm, S))
    assert abs(torch.stack(losses).mean().item() - lref.item()) < 1e-4 * max(1.0, abs(lref.item()))
    lse_all = torch.cat([torch.cat([l[:b] for l in lses]), torch.cat([l[b:] for l in lses])])
    for r in range(W):
        z, norm, S = locals_[r]
        dx = ops.ntxent_bwd(z, norm, zc, lse_all, S, b, r * b, b * W, 0.1, 1.0 / (2 * b * W))
        close(dx[:b], xr0.grad[r * b:(r + 1) * b], tol=1e-3)
        close(dx[b:], xr1.grad[r * b:(r + 1) * b], tol=1e-3)


def test_ntxent_packed_rank_major_layout_matches_single_process():
    """The layout loss.py uses across ranks: every 'rank' contributes ONE packed block [term0: out0,out1 | term1: out0,out1]
    to the gathered buffer, logits columns are found through the column map (blk = 2b, ld = 4b, base = term * 2b), the
    row log-sum-exps are gathered in the same packed layout.  Both terms must reproduce the single-process loss/gradient."""
    from vipformer_b200 import ops
    from vipformer_b200.loss import shard_layout

    b, W, D = 24, 3, 256
    xs = [(rnd((b * W, D), 10 + 2 * t), rnd((b * W, D), 11 + 2 * t)) for t in range(2)]
    refs = []
    for x0, x1 in xs:
        r0, r1 = x0.clone().requires_grad_(True), x1.clone().requires_grad_(True)
        l = ntxent_ref(r0, r1, 0.1)
        l.backward()
        refs.append((l.item(), r0.grad, r1.grad))
    n_r, n = 2 * b, 4 * b
    packs = []
    for r in range(W):
        sl = slice(r * b, (r + 1) * b)
        xp = torch.cat([xs[0][0][sl], xs[0][1][sl], xs[1][0][sl], xs[1][1][sl]], 0).contiguous()
        packs.append(ops.l2norm_rows(xp))
    zc = torch.cat([z for z, _ in packs], 0).contiguous()           # what all_gather_into_tensor produces
    lse_all = torch.empty(W * n, device="cuda")
    keep = []
    for r in range(W):
        z, norm = packs[r]
        off, half = shard_layout(r, W, b)
        for t in range(2):
            loss = torch.zeros(1, device="cuda")
            lse, S = ops.ntxent_fwd(z[t * n_r:(t + 1) * n_r], zc, b, off, half, 0.1, loss, n_c=W * n_r, colmap=(n_r, n, t * n_r),
                                    lse_out=lse_all[r * n + t * n_r:r * n + (t + 1) * n_r])
            keep.append((r, t, loss, S))
    for t in range(2):
        mean = torch.stack([l for r, tt, l, S in keep if tt == t]).mean().item()
        assert abs(mean - refs[t][0]) < 1e-4 * max(1.0, abs(refs[t][0]))
    for r, t, loss, S in keep:
        z, norm = packs[r]
        off, half = shard_layout(r, W, b)
        dx = ops.ntxent_bwd(z[t * n_r:(t + 1) * n_r], norm[t * n_r:(t + 1) * n_r], zc, lse_all, S, b, off, half, 0.1,
                            1.0 / (2 * b * W), n_c=W * n_r, colmap=(n_r, n, t * n_r))
        close(dx[:b], refs[t][1][r * b:(r + 1) * b], tol=1e-3)
        close(dx[b:], refs[t][2][r * b:(r + 1) * b], tol=1e-3)


def test_adamw_matches_torch():
    from vipformer_b200 import ops

    n = 100003
    p0, g = rnd(n, 1), rnd(n, 2)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=BF16)
    state = torch.tensor([0, 123], device="cuda", dtype=torch.int64)
    lr = torch.tensor([1e-3], device="cuda")
    for it in range(3):
        ref.grad = g * (it + 1)
        opt.step()
        ops.step_advance(state)
        ops.adamw(p, g * (it + 1), m, v, shadow, lr, state[:1])
    close(p, ref.detach(), tol=1e-5)
    assert torch.equal(shadow.float(), p.to(BF16).float())
    assert state[0].item() == 3 and state[1].item() != 123


def test_gelu_streaming_kernels():
    from vipformer_b200 import ops

    z = rnd((3000, 512), 1, BF16, scale=2.0)
    h = ops.gelu_fwd(z)
    close(h, F.gelu(z.float()), bf16=True)
    dh = rnd((3000, 512), 2, BF16)
    cs = torch.zeros(512, device="cuda")
    dz = ops.gelu_bwd(dh, z, colsum=cs)
    zf = z.float().requires_grad_(True)
    F.gelu(zf).backward(dh.float())
    close(dz, zf.grad, bf16=True)
    close(cs, dz.float().sum(0), tol=2e-3)
