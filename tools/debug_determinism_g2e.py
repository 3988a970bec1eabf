"""Which kernel of Group2Emb's forward is not bitwise reproducible?  Records every ops.* output of repeated forwards."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import _synth
from vipformer_b200 import ops
from vipformer_b200.model.pointcloud.utils import divide_patches

cfg = _synth.MODEL_CASES[sys.argv[1] if len(sys.argv) > 1 else "small"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.manual_seed(0)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
pc = pc.cuda().train()
pts, _, _ = _synth.model_inputs(cfg)
pts = pts.cuda().float().contiguous()
start = torch.arange(pts.shape[0], device="cuda") % cfg["N"]
nb, ce = divide_patches(pts, pc.num_groups, pc.group_size, start_idx=start)

log = []


def flat(o):
    if torch.is_tensor(o):
        return [o]
    if isinstance(o, (tuple, list)):
        return [t for x in o for t in flat(x)]
    if hasattr(o, "__dict__"):
        return [t for x in vars(o).values() for t in flat(x)]
    return []


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **k):
        r = orig(*a, **k)
        outs = flat(r) + [t for t in flat(list(a[2:3])) if t is not None] + \
            [k[x] for x in ("gm_bf16", "gm_f32", "gm_argmax") if k.get(x) is not None]
        log.append((name, [t.clone() for t in outs]))
        return r
    setattr(ops, name, f)


for n in ("linear3_stats", "bn_stats_finalize", "linear3_fwd", "gemm", "bn_forward"):
    wrap(n)


def run():
    log.clear()
    with torch.no_grad():
        pc.group2emb(nb)
    torch.cuda.synchronize()
    return list(log)


ref = run()
count = [0] * len(ref)
for _ in range(reps):
    cur = run()
    for i, ((n, a), (_, b)) in enumerate(zip(ref, cur)):
        if any(not torch.equal(x, y) for x, y in zip(a, b)):
            count[i] += 1
for i, (n, a) in enumerate(ref):
    print(i, n, [tuple(t.shape) for t in a], "differs in", count[i], "of", reps)
