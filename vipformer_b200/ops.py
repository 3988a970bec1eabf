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
                ("resid", _vp), ("seed_ptr", _vp)]


def _dp(t):
    return None if t is None else t.data_ptr()


def gemm(a, b, out, *, a_mn=False, b_mn=False, M=None, N=None, K=None, lda=None, ldb=None, mode=EPI_STORE,
         bias=None, rg_bias=None, rg_shift=0, act=ACT_NONE, aux=None, aux_mode=AUX_NONE, out2=None, resid=None,
         out_bf16=None, alpha=1.0, drop_p=0.0, seed=None, op_id=0, splits=None, ldc=None):
    """out (op)= epilogue(alpha * A @ B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn); bf16."""
    _lib.require_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    ldc = out.stride(0) if ldc is None else ldc
    e = GemmEpilogue()
    e.mode, e.out_f32, e.ldc, e.act = mode, int(out.dtype == torch.float32), ldc, act
    if mode != EPI_STORE:
        assert out.dtype == torch.float32
    else:
        assert out.dtype in (torch.float32, torch.bfloat16)
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
