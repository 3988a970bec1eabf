"""Cycles per epilogue phase of the tcgen05 GEMM (needs a build with VPF_NVCC_EXTRA=-DVPF_GEMM_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from vipformer_b200 import ops, _lib

NAMES = ["0 setup", "1 wait_read + ld issue", "2 wait tmem_full", "3 tcgen05.ld+wait", "4 tmem_empty arrive", "5 math+staging",
         "6 fence.proxy.async", "7 syncwarp + store issue", "8 pool / loop end", "9 -", "10 -"]
lib = _lib.lib() if hasattr(_lib, "lib") else _lib._LIB
M = 148 * 128 * 8
for out_dtype, N, K in ((torch.bfloat16, 512, 64), (torch.bfloat16, 512, 256), (torch.float32, 512, 256), (torch.bfloat16, 512, 1024)):
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((N, K), device="cuda").to(torch.bfloat16)
    out = torch.empty((M, N), device="cuda", dtype=out_dtype)
    bias = torch.randn(N, device="cuda")
    buf = (ctypes.c_ulonglong * 16)()
    for _ in range(2):
        ops.gemm(A, B, out, bias=bias)
    lib.vpf_debug_gemm_phase(buf)
    ops.gemm(A, B, out, bias=bias)
    lib.vpf_debug_gemm_phase(buf)
    n = max(1, buf[12])
    tot = sum(buf[i] for i in range(11))
    print(f"--- {out_dtype} N={N} K={K}: {n} half-steps, {tot / n:.0f} cycles per half-step")
    for i, nm in enumerate(NAMES):
        print(f"   {nm:28s} {buf[i] / n:8.0f} cyc  {100.0 * buf[i] / tot:5.1f}%")
