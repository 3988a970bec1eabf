"""Host-side cost of one eager step (what bounds world_size > 1, where the step is not graph-replayed)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import _synth, bench
from vipformer_b200.engine import PretrainEngine

b = 256
cfg = dict(bench.CFG, b=b)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"], use_cuda_graph=False)
for _ in range(3):
    eng.step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    eng.step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts.append(((t1 - t) * 1e3, (time.perf_counter() - t) * 1e3))
print("host issue ms / step wall ms:", " ".join("%.1f/%.1f" % x for x in ts))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    eng.step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:4500])
