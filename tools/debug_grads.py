import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth
from oracle import model_ref as M
from test_oracle_model_golden import oracle_run
from vipformer_b200.model.pointcloud.utils import Group2Emb
from vipformer_b200.loss import pretrain_loss

rel = lambda a, b: ((a.detach().cpu().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30)).item()
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.detach().cpu().double().reshape(1, -1), b.detach().double().reshape(1, -1)).item()

print("=== Group2Emb S=1")
torch.manual_seed(0)
g2e = Group2Emb(256)
sd = _synth.perturb_state_dict(g2e.state_dict(), 5)
g2e.load_state_dict(sd)
gen = torch.Generator().manual_seed(1)
nb = torch.randn((6, 512, 1, 3), generator=gen) * 0.3
sdr = {"g." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
tok_ref = M.group2emb(sdr, "g", nb, True)
dtok = torch.randn(tok_ref.shape, generator=gen)
(tok_ref * dtok).sum().backward()
g2e = g2e.cuda().train()
tok = g2e(nb.cuda())
(tok * dtok.cuda()).sum().backward()
print("tok", rel(tok, tok_ref))
for k, p in g2e.named_parameters():
    r = sdr["g." + k].grad
    print(f"   {k:28s} rel {rel(p.grad, r):.4f} cos {cos(p.grad, r):.5f} |ref| {r.norm().item():.3e}")

for name in sys.argv[1:] or ["small"]:
    print("=== model", name)
    nm, _, bb = name.partition(":")
    cfg = dict(_synth.MODEL_CASES[nm])
    if bb:
        cfg["b"] = int(bb)
    o = oracle_run(cfg)
    pc, im = _synth.build_models(cfg)
    pc.load_state_dict({k: v.detach() for k, v in o["sd_pc"].items() if k in pc.state_dict()})
    im.load_state_dict({k: v.detach() for k, v in o["sd_im"].items() if k in im.state_dict()})
    pc, im = pc.cuda().train(), im.cuda().train()
    pts, start, imgs = o["inputs"]
    pc.fps_start_idx = torch.from_numpy(start).cuda()
    pf, pb = pc(pts.cuda()); jf, jb = im(imgs.cuda())
    print("pc_back", rel(pb, o["pc_back"]), "pc_feats", rel(pf, o["pc_feats"]), "im_back", rel(jb, o["im_back"]), "im_feats", rel(jf, o["im_feats"]))
    L = pretrain_loss(pf, jf)
    print("loss", L.detach().cpu().numpy(), o["loss"])
    L[0].backward()
    for tag, model, sdd in (("pc", pc, o["sd_pc"]), ("img", im, o["sd_im"])):
        for k, p in model.named_parameters():
            r = sdd[k].grad
            e = rel(p.grad, r)
            flag = " <<<<" if e > 0.06 and r.norm().item() > 1e-3 else ""
            print(f"   {tag} {k:60s} rel {e:.4f} cos {cos(p.grad, r):.5f} |ref| {r.norm().item():.3e}{flag}")
