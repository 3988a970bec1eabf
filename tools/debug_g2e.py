import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth
from oracle import model_ref as M, tokenizer as T
from vipformer_b200.model.pointcloud.utils import Group2Emb
import torch.nn.functional as F

for S, D, G, N, B in ((8, 128, 32, 128, 12), (32, 256, 128, 2048, 8)):
    torch.manual_seed(0)
    g2e = Group2Emb(D)
    sd = _synth.perturb_state_dict(g2e.state_dict(), 5)
    g2e.load_state_dict(sd)
    pts = _synth.make_clouds("randn", B, N, 3); start = _synth.make_start(B, N, 3)
    nb, ce = T.divide_patches(pts, G, S, start)
    nb = torch.from_numpy(nb)
    sdr = {"g." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    tok_ref = M.group2emb(sdr, "g", nb, True)
    gen = torch.Generator().manual_seed(1)
    dtok = torch.randn(tok_ref.shape, generator=gen)
    (tok_ref * dtok).sum().backward()
    g2e = g2e.cuda().train()
    tok = g2e(nb.cuda())
    (tok * dtok.cuda()).sum().backward()
    rel = lambda a, b: ((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm()).item()
    print(f"S={S} D={D}: tok rel {rel(tok, tok_ref):.4f}")
    for k, p in g2e.named_parameters():
        r = sdr["g." + k].grad
        print(f"   {k:28s} rel {rel(p.grad, r):.4f}  |ref| {r.norm().item():.3e}")
    # argmax flips of the final pool when the pre-pool activations are rounded to bf16
    with torch.no_grad():
        x = nb.reshape(-1, 3)
        x = F.linear(x, sd["first_conv.0.weight"][:, :, 0], sd["first_conv.0.bias"])
        x = F.relu(M._bn({"b." + k: v for k, v in sd.items()}, "b.first_conv.1", x, True))
        x = F.linear(x, sd["first_conv.3.weight"][:, :, 0], sd["first_conv.3.bias"])
        R = x.shape[0]
        a32 = x.view(-1, S, 128).argmax(1); a16 = x.bfloat16().float().view(-1, S, 128).argmax(1)
        print("   conv2 pool argmax flips under bf16 rounding:", (a32 != a16).float().mean().item())
