"""CUDA-event time of the attention core (forward / backward) at the step's shapes: pc self-attention (128 x 128),
pc cross-attention (128 x 2048), image (144 x 144); B = clouds (512) or images (256), 4 heads.  L2 flushed between runs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vipformer_b200 import ops

BF16 = torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(os.environ.get("REPS", "5"))
cases = [("sa", 512, 4, 128, 128), ("ca", 512, 4, 128, 2048), ("img", 256, 4, 144, 144)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
seed = torch.tensor([77], device="cuda", dtype=torch.int64)
for name, B, H, Lq, Lk in cases:
    if which not in ("all", name):
        continue
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((B * Lq, D), device="cuda", generator=g).to(BF16)
    kv = torch.randn((B * Lk, 2 * D), device="cuda", generator=g).to(BF16)
    do = torch.randn((B * Lq, D), device="cuda", generator=g).to(BF16)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    k, v = kv[:, :D], kv[:, D:]
    tf, tb = [], []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        o, lse = ops.attention_fwd(q, k, v, B, H, Lq, Lk, 0.125, 0.1, seed, 5)
        e1.record()
        ops.attention_bwd(q, k, v, o, do, lse, dq, dkv[:, :D], dkv[:, D:], B, H, Lq, Lk, 0.125, 0.1, seed, 5)
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    fl = 4.0 * B * H * Lq * Lk * 64
    mf, mb = sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2]
    print(f"{name}: fwd {mf * 1e3:8.1f} us ({fl / mf / 1e9:7.1f} TFLOP/s)   bwd {mb * 1e3:8.1f} us ({2.5 * fl / mb / 1e9:7.1f} TFLOP/s)", flush=True)
