"""Timeline of the tcgen05 attention forward, CTA 0 (needs a build with VPF_NVCC_EXTRA=-DVPF_ATTN_TIMING).
Per block n (ns relative to block 0's first stamp): softmax warp: wait s_full begin / s_full seen / maxima exchanged /
P written (p_full arrive) / pv_full seen / item stored;  MMA warp: S issued / PV issued."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from vipformer_b200 import ops, _lib

BF16 = torch.bfloat16
B, H, Lq, Lk = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (512, 4, 128, 128)))
D = H * 64
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn((B * Lq, D), device="cuda", generator=g).to(BF16)
kv = torch.randn((B * Lk, 2 * D), device="cuda", generator=g).to(BF16)
seed = torch.tensor([77], device="cuda", dtype=torch.int64)
for _ in range(3):
    ops.attention_fwd(q, kv[:, :D], kv[:, D:], B, H, Lq, Lk, 0.125, 0.1, seed, 5)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 256)()
_lib.lib().vpf_debug_attn_stamps(buf)
t0 = min(x for x in buf if x)
names = ["wait_s", "s_full", "xchg", "p_full", "pv_full", "stored", "S_issue", "PV_issue"]
print("block " + " ".join(f"{n:>9s}" for n in names))
for n in range(12):
    print(f"{n:5d} " + " ".join(f"{(buf[n * 8 + k] - t0) if buf[n * 8 + k] else -1:9d}" for k in range(8)))
