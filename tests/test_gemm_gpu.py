"""GPU: tcgen05 GEMM vs a plain PyTorch fp32 reference of the same contraction (bf16 inputs, fp32 accumulate).
Tolerance: inputs are identical bf16 values, so only accumulation order differs -> rel 1e-3 of the row scale for
fp32 outputs; bf16 outputs add one rounding (2^-8 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g)).to(torch.bfloat16)


def _close(out, ref, bf16_out):
    out, ref = out.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    tol = (1.0 / 128 if bf16_out else 2e-3) * scale
    err = (out - ref).abs().max().item()
    assert err <= tol, f"max err {err} > {tol} (scale {scale})"


SHAPES = [(128, 128, 64), (256, 128, 256), (384, 256, 512), (1000, 200, 136), (130, 72, 432), (4096, 256, 64),
          (96, 384, 384), (128, 64, 128), (20000, 256, 256)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_layouts(M, N, K, a_mn, b_mn, out_dtype):
    from vipformer_b200 import ops

    if (a_mn and M % 8) or (b_mn and N % 8) or K % 8:
        pytest.skip("row stride must be a multiple of 16 bytes")
    A = _mk((M, K), 1)
    Bm = _mk((N, K), 2)
    ref = A.float() @ Bm.float().t()
    a = A.t().contiguous() if a_mn else A
    b = Bm.t().contiguous() if b_mn else Bm
    out = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    _close(out, ref, out_dtype == torch.bfloat16)


@pytest.mark.parametrize("N", [264, 512])   # 512: the 128 x 256 tile variant (two 64-column halves per epilogue group)
def test_gemm_epilogues(N):
    from vipformer_b200 import ops

    M, K = 777, 320
    A, Bm = _mk((M, K), 3), _mk((N, K), 4)
    bias = torch.randn(N, device="cuda")
    acc = A.float() @ Bm.float().t()
    # bias + gelu, with pre-activation copy
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    ops.gemm(A, Bm, out, bias=bias, act=ops.ACT_GELU, out2=pre)
    _close(pre, acc + bias, True)
    _close(out, torch.nn.functional.gelu(acc + bias), True)
    # relu + alpha
    out = torch.empty((M, N), device="cuda", dtype=torch.float32)
    ops.gemm(A, Bm, out, bias=bias, act=ops.ACT_RELU, alpha=0.5)
    _close(out, torch.relu(0.5 * acc + bias), False)
    # row-group bias (rows >> 5)
    rg = torch.randn(((M + 31) // 32, N), device="cuda")
    ops.gemm(A, Bm, out, rg_bias=rg, rg_shift=5)
    _close(out, acc + rg.repeat_interleave(32, 0)[:M], False)
    # gelu-grad / relu-mask aux
    z = _mk((M, N), 5)
    ops.gemm(A, Bm, out, aux=z, aux_mode=ops.AUX_GELU_GRAD)
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).sum().backward()
    _close(out, acc * zf.grad, False)
    ops.gemm(A, Bm, out, aux=z, aux_mode=ops.AUX_RELU_MASK)
    _close(out, acc * (z.float() > 0), False)
    # residual without dropout
    resid = torch.randn((M, N), device="cuda")
    ops.gemm(A, Bm, out, bias=bias, mode=ops.EPI_RESIDUAL, resid=resid)
    _close(out, resid + acc + bias, False)


def test_gemm_residual_dropout_is_deterministic_and_unbiased():
    from vipformer_b200 import ops

    M, N, K = 2048, 256, 64
    A, Bm = _mk((M, K), 6), _mk((N, K), 7)
    acc = A.float() @ Bm.float().t()
    resid = torch.zeros((M, N), device="cuda")
    seed = torch.tensor([1234], device="cuda", dtype=torch.int64)
    o1 = torch.empty((M, N), device="cuda")
    o2 = torch.empty_like(o1)
    ops.gemm(A, Bm, o1, mode=ops.EPI_RESIDUAL, resid=resid, drop_p=0.5, seed=seed, op_id=3)
    ops.gemm(A, Bm, o2, mode=ops.EPI_RESIDUAL, resid=resid, drop_p=0.5, seed=seed, op_id=3)
    assert torch.equal(o1, o2)
    kept = o1 != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.5) < 0.01
    _close(o1[kept], 2.0 * acc[kept], False)
    ops.gemm(A, Bm, o2, mode=ops.EPI_RESIDUAL, resid=resid, drop_p=0.5, seed=seed, op_id=4)
    assert not torch.equal(o1, o2)


@pytest.mark.parametrize("M,N,K", [(256, 256, 65536), (512, 256, 20000), (64, 432, 9216), (128, 3 * 256, 4096)])
def test_gemm_splitk_weight_grad(M, N, K):
    """dW[M=N_out, N=K_in] = dY^T X: both operands MN-major, reduction over tokens, atomic accumulate."""
    from vipformer_b200 import ops

    dY = _mk((K, M), 8)   # [tokens, out_features]
    X = _mk((K, N), 9)    # [tokens, in_features]
    ref = dY.float().t() @ X.float()
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(dY, X, out, a_mn=True, b_mn=True, mode=ops.EPI_ATOMIC_ADD)
    torch.cuda.synchronize()
    _close(out, ref, False)
    ops.gemm(dY, X, out, a_mn=True, b_mn=True, mode=ops.EPI_ATOMIC_ADD)   # accumulates
    _close(out, 2 * ref, False)


def test_gemm_strided_views():
    """q/k/v slices of a fused [T, 3D] buffer are consumed in place through lda."""
    from vipformer_b200 import ops

    T, D = 640, 256
    qkv = _mk((T, 3 * D), 10)
    W = _mk((D, D), 11)
    out = torch.empty((T, D), device="cuda")
    v = qkv[:, 2 * D:]
    ops.gemm(v, W, out)
    _close(out, v.float() @ W.float().t(), False)


@pytest.mark.parametrize("S", [8, 16, 32])
@pytest.mark.parametrize("store", [True, False])
def test_gemm_group_max_epilogue(S, store):
    """Fused per-patch max (utils.py:180,188) on the fp32 accumulators; first maximal row wins."""
    from vipformer_b200 import ops

    G, N, K = 77, 256, 128
    M = G * S
    A, Bm = _mk((M, K), 12), _mk((N, K), 13)
    bias = torch.randn(N, device="cuda")
    acc = A.float() @ Bm.float().t() + bias
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16) if store else None
    gm_f = torch.empty((G, N), device="cuda")
    gm_b = torch.empty((G, N), device="cuda", dtype=torch.bfloat16)
    am = torch.empty((G, N), device="cuda", dtype=torch.uint8)
    ops.gemm(A, Bm, out, bias=bias, gm_S=S, gm_f32=gm_f, gm_bf16=gm_b, gm_argmax=am)
    ref, ri = acc.view(G, S, N).max(1)
    _close(gm_f, ref, False)
    _close(gm_b, ref, True)
    if store:
        _close(out, acc, True)
    # argmax must point at a row whose value equals the maximum up to accumulation-order noise
    picked = acc.view(G, S, N).gather(1, am.long()[:, None]).squeeze(1)
    assert (ref - picked).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert (am.long() == ri).float().mean().item() > 0.995


@pytest.mark.parametrize("S", [8, 32])
def test_gemm_group_max_transposed(S):
    """Pool-only epilogue with the patch along the columns (the GEMM is called as C^T = W . X^T)."""
    from vipformer_b200 import ops

    G, C, K = 301, 256, 192
    R = G * S
    X, W = _mk((R, K), 14), _mk((C, K), 15)
    bias = torch.randn(C, device="cuda")
    acc = X.float() @ W.float().t() + bias
    gm_f = torch.empty((G, C), device="cuda")
    am = torch.empty((G, C), device="cuda", dtype=torch.uint8)
    ops.gemm(W, X, None, row_bias=bias, gm_S=S, gm_f32=gm_f, gm_argmax=am, gm_cols=True)
    ref, ri = acc.view(G, S, C).max(1)
    _close(gm_f, ref, False)
    picked = acc.view(G, S, C).gather(1, am.long()[:, None]).squeeze(1)
    assert (ref - picked).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert (am.long() == ri).float().mean().item() > 0.995
