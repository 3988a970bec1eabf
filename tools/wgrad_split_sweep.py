"""Weight-gradient GEMM (dW[N_out, K_in] += dY^T X, split-K + reduce-add) time against the number of K splits."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vipformer_b200 import ops

BF16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
for (T, No, Ki) in [(65536, 256, 256), (65536, 768, 256), (65536, 512, 256), (65536, 256, 512), (36864, 256, 256), (36864, 768, 256)]:
    dy = torch.randn((T, No), device="cuda", generator=g).to(BF16)
    x = torch.randn((T, Ki), device="cuda", generator=g).to(BF16)
    dW = torch.zeros((No, Ki), device="cuda")
    out = []
    for splits in (0, 8, 16, 24, 37, 74):
        ts = []
        for i in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(dy, x, dW, a_mn=True, b_mn=True, mode=ops.EPI_ATOMIC_ADD, splits=splits)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        out.append(f"splits={splits or 'auto'}: {sorted(ts)[len(ts) // 2]:6.1f} us")
    mb = (T * No + T * Ki) * 2 / 1e6
    print(f"T={T} dW[{No},{Ki}]  operands {mb:6.1f} MB ({mb / 6542.7 * 1e3:5.1f} us at the copy peak)   " + "   ".join(out), flush=True)
