"""GPU: the product modules (hand-written kernels) against the oracle restatement and the reference goldens.

Tolerances (bf16 tensor-core operands, fp32 accumulate; floors measured on the reference itself are in BASELINE.md 4):
embeddings rel-Frobenius <= 2e-2, |loss - oracle| <= 5e-2, per-parameter gradients rel-Frobenius <= 6e-2 for every
parameter whose gradient is not analytically zero (biases in front of a train-mode BatchNorm)."""
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
    assert relfro(pc_feats, o["pc_feats"]) < 3e-2 and relfro(pc_feats, g["pc_feats"]) < 3e-2
    assert relfro(im_feats, o["im_feats"]) < 3e-2
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
            if r > 6e-2:
                bad.append((tag, k, r))
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
    assert torch.equal(a, b) and not torch.equal(a, c)
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
