import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth
from oracle import model_ref as M
from vipformer_b200.model.pointcloud.partseg import Encoder
rel = lambda a, b: ((a.detach().cpu().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30)).item()
for (D, H, L, Lk, B, nsa, scale) in [(256, 4, 128, 300, 3, 0, 1.0), (256, 4, 128, 320, 3, 0, 1.0), (256, 4, 128, 300, 1, 0, 1.0), (256, 4, 128, 300, 3, 2, 1.0),
                                      (256, 4, 128, 300, 3, 0, 0.3), (128, 2, 32, 128, 3, 0, 1.0), (256, 4, 128, 256, 3, 1, 1.0)]:
    torch.manual_seed(3)
    enc = Encoder(num_latent_channels=D, num_cross_attention_heads=H, cross_attention_widening_factor=2, num_self_attention_layers=nsa,
                  num_self_attention_heads=H, self_attention_widening_factor=2, dpr_list=[0.0] * nsa, modal_prior=True)
    sd = _synth.perturb_state_dict(enc.state_dict(), 7)
    enc.load_state_dict(sd)
    gen = torch.Generator().manual_seed(4)
    x, pos, kv = (torch.randn(s, generator=gen) * scale for s in ((B, L, D), (B, L, D), (B, Lk, D)))
    sdr = {"e." + k: v.clone() for k, v in sd.items()}
    for k in list(sdr):
        if ".cross_attn_n." in k:
            sdr[k.replace(".cross_attn_n.", ".cross_attn_1.")] = sdr[k]
    with torch.no_grad():
        yr = M.encoder(sdr, "e", x, pos, kv, H, nsa)
        # stage-wise reference of the CA layer
        a = "e.cross_attn_1.0.module"
        xq = x + pos
        att = M.mha(sdr, a + ".attention", M._ln(sdr, a + ".q_norm", xq), M._ln(sdr, a + ".kv_norm", kv), H)
    enc = enc.cuda().eval()
    with torch.no_grad():
        y = enc(x.cuda(), pos.cuda(), kv.cuda())
    print(f"D={D} H={H} L={L} Lk={Lk} B={B} nsa={nsa} scale={scale}: out rel {rel(y, yr):.4f}  |att|/|xq| {att.norm().item() / xq.norm().item():.3f} max|logit-ish| {att.abs().max().item():.2f}")
