"""One eager pre-training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth
from vipformer_b200.engine import PretrainEngine
sys.path.insert(0, ROOT)
import bench

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(bench.CFG, b=b)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"], use_cuda_graph=False,
                     overlap_branches=bool(int(os.environ.get("VPF_PROFILE_OVERLAP", "0"))))
g = torch.Generator(device="cuda").manual_seed(1)
eng.pc_in.copy_(torch.randn(eng.pc_in.shape, device="cuda", generator=g) * 0.3)
eng.img_in.copy_(torch.randn(eng.img_in.shape, device="cuda", generator=g))
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", eng.losses.tolist())
