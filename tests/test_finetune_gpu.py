"""GPU: CrossFormer_pc_mp_ft + label-smoothed cross entropy (BASELINE configs[3]) against the oracle,
which tests/test_oracle_model_golden.py pins to vectors from the real reference (tests/golden/model_ft_*.npz)."""
import os

import numpy as np
import pytest
import torch

import _synth
from test_oracle_model_golden import oracle_ft_run

pytestmark = pytest.mark.gpu


def relfro(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _pins(tap):
    g, a, h = tap["g2e"][0], tap["adapter"][0], tap["cls_head"][0]
    c = lambda t: t.detach().cpu()
    return {"pc.g2e.relu1": c(g.h1 > 0), "pc.g2e.max2": c(g.am2).long(), "pc.g2e.relu3": c(g.h3 > 0), "pc.g2e.max4": c(g.am4).long(),
            "pc.adapter.relu": c(a.h > 0), "pc.pool.max": c(h.am).long(), "pc.cls.relu1": c(h.a[0] > 0),
            "pc.cls.relu2": c(h.a[1] > 0), "pc.cls.relu3": c(h.a[2] > 0)}


@pytest.mark.parametrize("name", ["ft_small", "ft_cfgA"])
def test_finetune_forward_loss_backward_match_oracle(name, golden_dir):
    import vipformer_b200.runtime as rt
    from vipformer_b200.loss import CrossEntropyLoss

    cfg = _synth.FT_CASES[name]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    o0 = oracle_ft_run(cfg)
    model = _synth.build_ft_model(cfg)
    model.load_state_dict({k: v.detach() for k, v in o0["sd"].items() if k in model.state_dict()})
    model = model.cuda().train()
    pts, start, labels = o0["inputs"]
    model.fps_start_idx = torch.from_numpy(start).cuda()
    rt.TAP = {}
    try:
        logits = model(pts.cuda())
        tap = rt.TAP
    finally:
        rt.TAP = None
    assert logits.shape == (cfg["b"], cfg["classes"]) and logits.dtype == torch.float32
    loss = CrossEntropyLoss(label_smoothing=0.2)(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    o = oracle_ft_run(cfg, pins=_pins(tap))
    # logits sit behind three train-mode BatchNorms over 8-10 samples: 5e-2 (like the projected features of the pre-train models)
    assert relfro(logits, o["logits"]) < 5e-2 and relfro(logits, torch.from_numpy(g["logits"])) < 8e-2
    assert abs(loss.item() - o["loss"]) < 5e-2 and abs(loss.item() - float(g["loss"][0])) < 5e-2
    gmax = max(o["sd"][k].grad.norm().item() for k in o["names"] if o["sd"][k].grad is not None)
    bad, worst = [], 0.0
    for k, p in model.named_parameters():
        ref = o["sd"][k].grad
        if ref is None or ref.norm().item() < 1e-4 * gmax:      # unused latent_head / biases in front of a train-mode BN
            assert p.grad is None or p.grad.float().norm().item() <= 1e-2 * gmax, k
            continue
        e = relfro(p.grad, ref)
        worst = max(worst, e)
        if e > 1e-1:
            bad.append((k, round(e, 4)))
    print(f"[{name}] fine-tune: worst per-parameter rel-Frobenius gradient error with pinned choices {worst:.4f}")
    assert not bad, bad
    sdm = model.state_dict()
    for k, v in o["run"].items():
        assert relfro(sdm[k], v) < 2e-2, k


def test_pretrained_checkpoint_loads_into_finetune_model():
    """ft_cls.py:92-98: keys of a (DDP-saved) pre-training checkpoint load with strict=False; only finetune_head.* is missing."""
    from vipformer_b200.model.pointcloud import load_pretrained

    cfg = _synth.FT_CASES["ft_small"]
    pc, _ = _synth.build_models(dict(cfg, img=144, patch=12))
    ck = {"module." + k: v + 0.5 for k, v in pc.state_dict().items() if v.dtype.is_floating_point}
    ft = _synth.build_ft_model(cfg)
    missing, unexpected = load_pretrained(ft, ck)
    assert not unexpected and missing and all(k.startswith("finetune_head.") or "num_batches_tracked" in k for k in missing)
    assert torch.equal(ft.state_dict()["group2emb.first_conv.0.weight"], ck["module.group2emb.first_conv.0.weight"])
    ft = ft.cuda().train()
    pts, start, labels = _synth.ft_inputs(cfg)
    ft.fps_start_idx = torch.from_numpy(start).cuda()
    assert torch.isfinite(ft(pts.cuda())).all()


@pytest.mark.parametrize("n,C,eps", [(10, 15, 0.2), (64, 40, 0.2), (7, 50, 0.0)])
def test_cross_entropy_label_smoothing_kernel(n, C, eps):
    from vipformer_b200.loss import CrossEntropyLoss

    g = torch.Generator(device="cuda").manual_seed(n)
    x = (torch.randn((n, C), device="cuda", generator=g) * 3).requires_grad_(True)
    y = torch.randint(0, C, (n,), device="cuda", generator=g)
    loss = CrossEntropyLoss(label_smoothing=eps)(x, y)
    (2.5 * loss).backward()
    xr = x.detach().clone().requires_grad_(True)
    lr = torch.nn.functional.cross_entropy(xr, y, label_smoothing=eps)
    (2.5 * lr).backward()
    assert abs(loss.item() - lr.item()) < 1e-5 * max(1.0, abs(lr.item()))
    assert (x.grad - xr.grad).abs().max().item() < 1e-6
