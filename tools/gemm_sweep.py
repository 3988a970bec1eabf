"""Time per tile of the tcgen05 GEMM against K (slope = per-k-block cost, intercept = per-tile epilogue cost)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from vipformer_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
dev = "cuda"
for out_dtype in (torch.bfloat16, torch.float32):
    for N in (128, 256, 512):
        for K in (64, 128, 256, 512, 1024, 2048):
            A = torch.randn((M, K), device=dev).to(torch.bfloat16)
            B = torch.randn((N, K), device=dev).to(torch.bfloat16)
            out = torch.empty((M, N), device=dev, dtype=out_dtype)
            for _ in range(3):
                ops.gemm(A, B, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                ops.gemm(A, B, out)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            tiles128 = (M // 128) * (N // 128)
            per_sm = tiles128 / 148
            print(f"{str(out_dtype)[6:]:9s} M={M} N={N:4d} K={K:5d}  {us:8.1f} us  {us / per_sm:6.3f} us per 128x128 tile per SM  "
                  f"{2.0 * M * N * K / us / 1e6:7.1f} TF/s  fill {tiles128 * K * 512 / 148 / us / 1e3:6.1f} GB/s/SM")
