"""Repeat the forward pass with frozen weights / seed and report which intermediate first differs bitwise."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import _synth
from vipformer_b200.model.pointcloud.utils import divide_patches

cfg = _synth.MODEL_CASES[sys.argv[1] if len(sys.argv) > 1 else "small"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.manual_seed(0)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
pc, im = pc.cuda().train(), im.cuda().train()
pts, _, imgs = _synth.model_inputs(cfg)
pts, imgs = pts.cuda(), imgs.cuda()
start = torch.arange(pts.shape[0], device="cuda") % cfg["N"]
pc.fps_start_idx = start


def pc_stages():
    out = {}
    with torch.no_grad():
        p = pts.float().contiguous()
        out["pts_embs"] = pc.input_adapter(p)
        nb, ce = divide_patches(p, pc.num_groups, pc.group_size, start_idx=start)
        out["neigh"], out["center"] = nb, ce
        out["group_embs"] = pc.group2emb(nb)
        out["pos_embs"] = pc.position_emb(ce)
        x = pc.encoder(out["group_embs"], out["pos_embs"], out["pts_embs"])
        out["latent"] = x
        f, bk = pc.latent_head(x)
        out["feats"], out["backbone"] = f, bk
    return {k: v.clone() for k, v in out.items()}


def img_stages():
    out = {}
    with torch.no_grad():
        f, bk = im(imgs)
        out["img_feats"], out["img_backbone"] = f, bk
    return {k: v.clone() for k, v in out.items()}


for name, fn in (("pc", pc_stages), ("img", img_stages)):
    ref = fn()
    diffs = {k: 0 for k in ref}
    worst = {k: 0.0 for k in ref}
    for _ in range(reps):
        cur = fn()
        for k in ref:
            if not torch.equal(cur[k], ref[k]):
                diffs[k] += 1
                worst[k] = max(worst[k], float((cur[k].float() - ref[k].float()).abs().max()))
    print(name, {k: (diffs[k], worst[k]) for k in ref})
