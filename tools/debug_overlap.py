"""Losses / parameter drift of the two-stream schedule against the serial one (and serial against itself)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import _synth
from vipformer_b200.engine import PretrainEngine

cfg = _synth.MODEL_CASES[sys.argv[1] if len(sys.argv) > 1 else "small"]
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3


def run(overlap):
    torch.manual_seed(0)
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    eng = PretrainEngine(pc, im, batch_pairs=cfg["b"], num_points=cfg["N"], lr=lr, use_cuda_graph=graph, seed=11,
                         overlap_branches=overlap)
    pts, _, imgs = _synth.model_inputs(cfg)
    eng.pc_in.copy_(pts.cuda()); eng.img_in.copy_(imgs.cuda().permute(0, 3, 1, 2))
    hist, grads = [], []
    for _ in range(4):
        hist.append(eng.step().clone())
        grads.append(eng.arena.flat_g.clone())
    torch.cuda.synchronize()
    return torch.stack(hist).cpu(), grads, eng.arena.flat_p.clone()


res = [run(o) for o in (False, False, True, True)]
for name, (h, g, p) in zip(["serial-a", "serial-b", "overlap-a", "overlap-b"], res):
    print(name, ["%.6f" % v for v in h[:, 0].tolist()])
ref = res[0]
for name, (h, g, p) in zip(["serial-b", "overlap-a", "overlap-b"], res[1:]):
    print(name, "grad rel per step:", [float((a - b).norm() / b.norm()) for a, b in zip(g, ref[1])],
          "param rel:", float((p - ref[2]).norm() / ref[2].norm()))
