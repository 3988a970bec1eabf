"""Wall time per step of the synchronous and the prefetched host-input paths."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import _synth, bench
from vipformer_b200.engine import PretrainEngine

b = 256
cfg = dict(bench.CFG, b=b)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"])
host = [(torch.randn(b, cfg["N"], 3).pin_memory(), torch.randn(b, cfg["N"], 3).pin_memory(), torch.randn(b, 3, 144, 144).pin_memory()) for _ in range(2)]
for _ in range(4):
    eng.step_host(*host[0])
def run(mode, n=12):
    torch.cuda.synchronize(); ts = []
    if mode == "prefetch": eng.prefetch_host(*host[0])
    for i in range(n):
        t = time.perf_counter()
        if mode == "sync": eng.step_host(*host[i % 2])
        elif mode == "prefetch": eng.step_host_prefetched(host[(i + 1) % 2] if i + 1 < n else None)
        elif mode == "none": eng.step(); eng.losses_host.copy_(eng.losses, non_blocking=True); torch.cuda.current_stream().synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    print(mode, " ".join("%.2f" % x for x in ts))
for m in ("none", "sync", "prefetch", "none", "sync", "prefetch"):
    run(m)
# H2D alone
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    eng.pc_in[:b].copy_(host[0][0], non_blocking=True); eng.pc_in[b:].copy_(host[0][1], non_blocking=True); eng.img_in.copy_(host[0][2], non_blocking=True)
torch.cuda.synchronize(); print("h2d alone ms", (time.perf_counter() - t) * 1e3 / 5)
