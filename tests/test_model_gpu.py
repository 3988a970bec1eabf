"""GPU: the product modules (hand-written kernels) against the oracle restatement and the reference goldens.

Tolerances (bf16 tensor-core operands, fp32 accumulate; floors measured on the reference itself are in BASELINE.md 4):
backbone embeddings rel-Frobenius <= 2e-2, projected features <= 5e-2 (they sit behind two train-mode BatchNorms over
only 6..12 samples in these fixtures, which amplifies the backbone error), |loss - oracle| <= 5e-2.

Gradients.  Every block is checked against the oracle's autograd with a GIVEN upstream gradient and exact inputs
(test_encoder_block_backward <= 3e-2; test_pool_head_block_backward <= 8e-2 behind its two BatchNorms;
test_group2emb_backward_exact <= 2e-2, <= 0.12 behind its ReLU masks).  At MODEL level the comparison additionally contains DISCRETE choices that
the bf16 rounding of the forward activations can flip: the token max pool over 128 tokens (partseg.py:547), the two
per-patch max pools (utils.py:180,188) and ReLU masks at |pre-activation| ~ 0.  A 0.3 % forward perturbation flips
5-20 % of the 128-way token arg-maxes (tools/debug_grads.py: the error jumps from 13 % to 26 % exactly across the pool),
each flip moves a whole gradient row, and NT-Xent at T = 0.1 plus train-mode BatchNorm over 8..12 samples amplify the
rest -- although every kernel is exact and the forward features agree to 0.3 % (the reference under fp16 autocast has
the same effect at a lower rate).  Stated model-level gate: cosine >= 0.90 and rel-Frobenius <= 0.5 per parameter
tensor; parameters whose gradient is analytically zero (biases in front of a train-mode BatchNorm) are checked absolutely."""
import os

import numpy as np
import pytest
import torch

import _synth
from test_oracle_model_golden import oracle_run

pytestmark = pytest.mark.gpu


def relfro(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1) if not torch.is_tensor(b) else b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _load(models, o):
    pc, im = models
    pc.load_state_dict({k: v.detach() for k, v in o["sd_pc"].items() if k in pc.state_dict()})
    im.load_state_dict({k: v.detach() for k, v in o["sd_im"].items() if k in im.state_dict()})
    return pc.cuda().train(), im.cuda().train()


@pytest.fixture(scope="module")
def runs():
    out = {}
    for name in ("small", "cfgA"):
        cfg = _synth.MODEL_CASES[name]
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        o = oracle_run(cfg)
        pc, im = _load(_synth.build_models(cfg), o)
        pts, start, imgs = o["inputs"]
        pc.fps_start_idx = torch.from_numpy(start).cuda()
        out[name] = (cfg, o, pc, im, pts.cuda(), imgs.cuda())
    return out


@pytest.mark.parametrize("name", ["small", "cfgA"])
def test_forward_loss_backward_match_oracle(name, runs, golden_dir):
    from vipformer_b200.loss import pretrain_loss

    cfg, o, pc, im, pts, imgs = runs[name]
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    pc.zero_grad(set_to_none=True)
    im.zero_grad(set_to_none=True)
    pc_feats, pc_back = pc(pts)
    im_feats, im_back = im(imgs)
    assert pc_feats.dtype == torch.float32 and pc_feats.shape == (2 * cfg["b"], cfg["D"])
    assert relfro(pc_back, o["pc_back"]) < 2e-2 and relfro(pc_back, g["pc_backbone"]) < 2e-2
    assert relfro(im_back, o["im_back"]) < 2e-2 and relfro(im_back, g["img_backbone"]) < 2e-2
    assert relfro(pc_feats, o["pc_feats"]) < 5e-2 and relfro(pc_feats, g["pc_feats"]) < 5e-2
    assert relfro(im_feats, o["im_feats"]) < 5e-2
    losses = pretrain_loss(pc_feats, im_feats, temperature=0.1, cmid_weight=1.0)
    lv = losses.detach().cpu().numpy()
    assert np.all(np.abs(lv - np.array(o["loss"])) <= 5e-2), (lv, o["loss"])
    assert np.all(np.abs(lv - g["loss"]) <= 5e-2)
    losses[0].backward()
    torch.cuda.synchronize()
    bad = []
    for tag, model, sd, names in (("pc", pc, o["sd_pc"], o["pnames"]), ("img", im, o["sd_im"], o["inames"])):
        gmax = max(sd[k].grad.norm().item() for k in names)
        for k, p in model.named_parameters():
            ref = sd[k].grad
            assert p.grad is not None, k
            if ref.norm().item() < 1e-4 * gmax:      # analytically-zero gradients: absolute check
                if p.grad.float().norm().item() > 1e-2 * gmax:
                    bad.append((tag, k, "nonzero", p.grad.norm().item()))
                continue
            r = relfro(p.grad, ref)
            cos = torch.nn.functional.cosine_similarity(p.grad.detach().double().cpu().reshape(1, -1), ref.double().reshape(1, -1)).item()
            if r > 0.5 or cos < 0.90:
                bad.append((tag, k, r, cos))
    assert not bad, bad
    # running statistics (checkpoint parity): momentum 0.1, unbiased variance
    for tag, model, run in (("pc", pc, o["run_pc"]), ("img", im, o["run_im"])):
        sdm = model.state_dict()
        for k, v in run.items():
            assert relfro(sdm[k], v) < 2e-2, k


def test_eval_mode_is_deterministic_and_uses_running_stats(runs):
    cfg, o, pc, im, pts, imgs = runs["small"]
    pc.eval(); im.eval()
    with torch.no_grad():
        a, _ = pc(pts)
        b, _ = pc(pts)
        c, _ = im(imgs)
    pc.train(); im.train()
    assert torch.equal(a, b)
    assert torch.isfinite(a).all() and torch.isfinite(c).all()


def test_dropout_training_step_changes_with_seed(runs):
    import vipformer_b200.runtime as rt

    cfg = _synth.MODEL_CASES["small"]
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc = pc.cuda().train()
    pts = runs["small"][4]
    pc.fps_start_idx = runs["small"][2].fps_start_idx
    rt.manual_seed(1)
    a, _ = pc(pts)
    rt.manual_seed(1)
    b, _ = pc(pts)
    rt.manual_seed(2)
    c, _ = pc(pts)
    # same seed => same masks (outputs equal up to the last-bit nondeterminism of the atomically reduced BatchNorm
    # statistics); a different seed draws different masks
    same, diff = relfro(a, b), relfro(a, c)
    assert same < 0.05 and diff > 0.2 and diff > 4 * same, (same, diff)
    a.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in pc.parameters())


def test_submodules_callable_like_the_reference(runs):
    from vipformer_b200.model.pointcloud.partseg import Encoder, MultiHeadAttention
    from vipformer_b200.model.pointcloud.utils import Group2Emb

    enc = Encoder(num_latent_channels=128, num_cross_attention_heads=2, cross_attention_widening_factor=2,
                  num_self_attention_layers=1, num_self_attention_heads=2, self_attention_widening_factor=2,
                  dpr_list=[0.0], modal_prior=True).cuda()
    x = torch.randn(2, 32, 128, device="cuda", requires_grad=True)
    pos = torch.randn(2, 32, 128, device="cuda", requires_grad=True)
    kv = torch.randn(2, 100, 128, device="cuda", requires_grad=True)
    y = enc(x, pos, kv)
    y.square().mean().backward()
    assert x.grad is not None and pos.grad is not None and kv.grad is not None
    assert set(k.split(".")[0] for k in enc.state_dict()) == {"cross_attn_n", "cross_attn_1", "sa_layers"}
    g2e = Group2Emb(128).cuda()
    t = g2e(torch.randn(2, 8, 16, 3, device="cuda"))
    assert t.shape == (2, 8, 128)
    with pytest.raises(ValueError):
        MultiHeadAttention(3, 128, 128, 128)
    with pytest.raises(ValueError):
        Encoder(num_latent_channels=128, num_cross_attention_layers=0)
    with pytest.raises(NotImplementedError):
        enc.cross_attn_1(x, kv, attn_mask=torch.ones(1))


def test_group2emb_backward_exact():
    """group_size = 1 makes both max pools trivial (no arg-max flips): Group2Emb forward/backward must then match the
    oracle's autograd to bf16 GEMM accuracy (rel-Frobenius 3e-2), which checks every backward kernel of the block."""
    from oracle import model_ref as M
    from vipformer_b200.model.pointcloud.utils import Group2Emb

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
    assert relfro(tok, tok_ref) < 1e-2
    gmax = max(v.grad.norm().item() for v in sdr.values() if getattr(v, "grad", None) is not None)
    for k, p in g2e.named_parameters():
        ref = sdr["g." + k].grad
        if ref.norm().item() < 1e-4 * gmax:
            assert p.grad.norm().item() < 1e-2 * gmax, k
        elif k in ("second_conv.3.weight", "second_conv.1.weight", "second_conv.3.bias"):
            assert relfro(p.grad, ref) < 2e-2, (k, relfro(p.grad, ref))     # no discrete choice upstream of these
        else:                                                                # downstream of the BN+ReLU mask (flips)
            assert relfro(p.grad, ref) < 0.12, (k, relfro(p.grad, ref))


def test_encoder_block_backward():
    """Cross-attention layer + 2 self-attention layers (all smooth ops) with a given upstream gradient: forward and
    every gradient (inputs, positional term, K/V source, all parameters) within rel-Frobenius 3e-2 of the oracle."""
    from oracle import model_ref as M
    from vipformer_b200.model.pointcloud.partseg import Encoder

    torch.manual_seed(3)
    D, H, L, Lk, B = 256, 4, 128, 300, 3
    enc = Encoder(num_latent_channels=D, num_cross_attention_heads=H, cross_attention_widening_factor=2,
                  num_self_attention_layers=2, num_self_attention_heads=H, self_attention_widening_factor=2,
                  dpr_list=[0.0, 0.0], modal_prior=True)
    sd = _synth.perturb_state_dict(enc.state_dict(), 7)
    enc.load_state_dict(sd)
    gen = torch.Generator().manual_seed(4)
    x, pos, kv, dout = (torch.randn(s, generator=gen) for s in ((B, L, D), (B, L, D), (B, Lk, D), (B, L, D)))
    sdr = {"e." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    for k in list(sdr):
        if ".cross_attn_n." in k:
            sdr[k.replace(".cross_attn_n.", ".cross_attn_1.")] = sdr[k]
    xr, pr, kr = (t.clone().requires_grad_(True) for t in (x, pos, kv))
    yr = M.encoder(sdr, "e", xr, pr, kr, H, 2)
    (yr * dout).sum().backward()
    enc = enc.cuda().train()
    xg, pg, kg = (t.cuda().requires_grad_(True) for t in (x, pos, kv))
    y = enc(xg, pg, kg)
    (y * dout.cuda()).sum().backward()
    assert relfro(y, yr) < 1e-2
    assert relfro(xg.grad, xr.grad) < 3e-2 and relfro(pg.grad, pr.grad) < 3e-2 and relfro(kg.grad, kr.grad) < 3e-2
    for k, p in enc.named_parameters():
        assert relfro(p.grad, sdr["e." + k].grad) < 3e-2, (k, relfro(p.grad, sdr["e." + k].grad))


def test_pool_head_block_backward():
    """Token max/mean pooling + latent_head (2 x BatchNorm1d train mode, ReLU, bias-free Linears) with exact fp32
    inputs and a given upstream gradient for both outputs: no forward noise => no arg-max flips => tight match."""
    from oracle import model_ref as M
    from vipformer_b200.model.pointcloud.partseg import _latent_head

    torch.manual_seed(5)
    B, L, D = 64, 128, 256
    head = _latent_head(D)
    sd = _synth.perturb_state_dict(head.state_dict(), 9)
    head.load_state_dict(sd)
    gen = torch.Generator().manual_seed(6)
    x = torch.randn((B, L, D), generator=gen)
    df, db = torch.randn((B, D), generator=gen), torch.randn((B, 2 * D), generator=gen) * 0.1
    sdr = {"h." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    back_r = torch.cat([xr.max(1)[0], xr.mean(1)], 1)
    feats_r = M.latent_head(sdr, "h", back_r, True)
    ((feats_r * df).sum() + (back_r * db).sum()).backward()
    head = head.cuda().train()
    xg = x.cuda().requires_grad_(True)
    feats, back = head(xg)
    ((feats * df.cuda()).sum() + (back * db.cuda()).sum()).backward()
    assert relfro(back, back_r) < 1e-5 and relfro(feats, feats_r) < 2e-2
    # gradients pass two BatchNorm backward stages as bf16 GEMM operands: 8e-2 (the bias-free Linears themselves 3e-2)
    assert relfro(xg.grad, xr.grad) < 8e-2
    for k, p in head.named_parameters():
        tol = 3e-2 if k in ("5.weight", "3.weight") else 8e-2
        assert relfro(p.grad, sdr["h." + k].grad) < tol, (k, relfro(p.grad, sdr["h." + k].grad))


def test_forward_bitwise_reproducible():
    """Same weights, inputs, FPS start and dropout seed -> bit-identical forward outputs (no order-dependent fp32
    reduction on the forward path; BatchNorm statistics combine in a fixed order inside a CTA and in fp64 across)."""
    cfg = _synth.MODEL_CASES["small"]
    torch.manual_seed(0)
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc, im = pc.cuda().train(), im.cuda().train()
    pts, _, imgs = _synth.model_inputs(cfg)
    pts, imgs = pts.cuda(), imgs.cuda()
    pc.fps_start_idx = torch.arange(pts.shape[0], device="cuda") % cfg["N"]
    import vipformer_b200.runtime as rt

    with torch.no_grad():
        rt.manual_seed(5)           # (also rewinds the per-forward dropout epoch of the drop-in path)
        ref = [t.clone() for t in (*pc(pts), *im(imgs))]
        for _ in range(10):
            rt.manual_seed(5)
            cur = (*pc(pts), *im(imgs))
            assert all(torch.equal(a, b) for a, b in zip(ref, cur))


# Shape envelope of SURVEY.md 8(b): the other published configurations (D 384 / 6 heads / MLP ratio 4, 96 groups,
# 1024 or 2500 points), reduced in depth so the CPU oracle finishes in seconds.  No golden file: the oracle is pinned to the
# reference by tests/test_oracle_model_golden.py on `small` and `cfgA`.  Gradients are compared with the product's discrete
# choices pinned in the oracle (tests/test_parity_pinned_gpu.py explains why and where the gates come from).
ENVELOPE = {
    "B_D384_H6_MR4": dict(D=384, H=6, n_sa=2, G=128, S=32, N=1024, MR=4, b=8, img=144, patch=12, seed=31),
    "C_G96_N1024": dict(D=256, H=4, n_sa=2, G=96, S=32, N=1024, MR=2, b=8, img=144, patch=12, seed=32),
    "N2500": dict(D=256, H=4, n_sa=1, G=128, S=32, N=2500, MR=2, b=8, img=144, patch=12, seed=33),
}


@pytest.mark.parametrize("name", sorted(ENVELOPE))
def test_shape_envelope_matches_oracle(name):
    from test_parity_pinned_gpu import compare_grads, pins_from_tap, run_product

    cfg = ENVELOPE[name]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    r = run_product(cfg)
    o = oracle_run(cfg, pins=pins_from_tap(r["tap"], cfg))
    assert r["pc_feats"].shape == (2 * cfg["b"], cfg["D"]) and r["im_back"].shape == (cfg["b"], 2 * cfg["D"])
    assert relfro(r["pc_back"], o["pc_back"]) < 2e-2 and relfro(r["im_back"], o["im_back"]) < 2e-2
    assert relfro(r["pc_feats"], o["pc_feats"]) < 5e-2 and relfro(r["im_feats"], o["im_feats"]) < 5e-2
    assert np.all(np.abs(r["losses"] - np.array(o["loss"])) <= 5e-2)
    bad, worst = compare_grads(r, o, 1.5e-1)
    print(f"[{name}] worst per-parameter rel-Frobenius gradient error (NT-Xent, pinned choices): {worst:.4f}")
    assert not bad, bad
