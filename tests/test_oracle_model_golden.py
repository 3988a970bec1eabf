"""CPU: pin oracle/model_ref.py (the floating-point oracle) against golden vectors made from the REAL reference.
Tolerance: both sides are fp32 CPU PyTorch evaluating the same formulae in (slightly) different op order -> rel 1e-4."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import _synth
from oracle import model_ref as M


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def linear_upstream(cfg):
    """Fixed pseudo-random upstream gradients for (pc_feats, pc_backbone, img_feats, img_backbone): the 'loss' is the
    linear functional sum <output, G>, which takes NT-Xent's T = 0.1 noise amplification out of a gradient comparison."""
    g = torch.Generator().manual_seed(cfg["seed"] + 99)
    b, D = cfg["b"], cfg["D"]
    return (torch.randn((2 * b, D), generator=g), 0.1 * torch.randn((2 * b, 2 * D), generator=g),
            torch.randn((b, D), generator=g), 0.1 * torch.randn((b, 2 * D), generator=g))


def oracle_run(cfg, dtype=torch.float32, pins=None, drop=None, start=None, linear=False):
    """Shared with the GPU tests: product-mirror init (identical to the reference's, asserted by make_golden_model.py)
    -> perturbed state_dicts -> oracle forward/loss/backward.
    pins: discrete choices to impose (oracle.model_ref.choices); drop: kwargs of oracle.model_ref.dropout."""
    import contextlib
    pc, im = _synth.build_models(cfg)
    sd_pc = {k: v.to(dtype) if v.dtype.is_floating_point else v for k, v in _synth.perturb_state_dict(pc.state_dict(), cfg["seed"] + 10).items()}
    sd_im = {k: v.to(dtype) if v.dtype.is_floating_point else v for k, v in _synth.perturb_state_dict(im.state_dict(), cfg["seed"] + 11).items()}
    pnames = [k for k, _ in pc.named_parameters()]
    inames = [k for k, _ in im.named_parameters()]
    for k in pnames:
        sd_pc[k] = sd_pc[k].clone().requires_grad_(True)
    for k in inames:
        sd_im[k] = sd_im[k].clone().requires_grad_(True)
    for sd in (sd_pc, sd_im):   # cross_attn_1 and cross_attn_n are ONE module in the reference (partseg.py:295-300)
        for k in list(sd.keys()):
            if ".cross_attn_n." in k:
                sd[k.replace(".cross_attn_n.", ".cross_attn_1.")] = sd[k]
    pts, start0, imgs = _synth.model_inputs(cfg)
    start = start0 if start is None else np.asarray(start)
    run_pc, run_im = {}, {}
    with M.choices(pins) as ch, (M.dropout(**drop) if drop else contextlib.nullcontext()):
        pc_feats, pc_back = M.pc_forward(sd_pc, pts.to(dtype), start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], True, run_pc)
        im_feats, im_back = M.img_forward(sd_im, imgs.to(dtype), cfg["patch"], cfg["H"], cfg["n_sa"], True, run_im)
        total, imid, cmid = M.pretrain_loss(pc_feats, im_feats)
        if linear:
            G = linear_upstream(cfg)
            sum((t * g.to(dtype)).sum() for t, g in zip((pc_feats, pc_back, im_feats, im_back), G)).backward()
        else:
            total.backward()
    return dict(rec=ch.rec, sd_pc=sd_pc, sd_im=sd_im, pnames=pnames, inames=inames, pc_feats=pc_feats.detach(), pc_back=pc_back.detach(),
                im_feats=im_feats.detach(), im_back=im_back.detach(), loss=(total.item(), imid.item(), cmid.item()),
                run_pc=run_pc, run_im=run_im, inputs=(pts, start, imgs))


@pytest.mark.parametrize("name", ["small", "cfgA"])
def test_model_oracle_matches_reference(name, golden_dir):
    cfg = _synth.MODEL_CASES[name]
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    torch.set_num_threads(8)
    o = oracle_run(cfg)
    assert _rel(o["pc_feats"], g["pc_feats"]) < 1e-4
    assert _rel(o["pc_back"], g["pc_backbone"]) < 1e-4
    assert _rel(o["im_feats"], g["img_feats"]) < 1e-4
    assert _rel(o["im_back"], g["img_backbone"]) < 1e-4
    assert np.allclose(o["loss"], g["loss"], rtol=1e-4, atol=1e-5)
    for tag, sd, names, run in (("pc", o["sd_pc"], o["pnames"], o["run_pc"]), ("img", o["sd_im"], o["inames"], o["run_im"])):
        assert list(g[f"{tag}_grad_names"]) == names
        norms = np.array([sd[k].grad.double().norm().item() for k in names])
        ref = g[f"{tag}_grad_norms"]
        # biases in front of a train-mode BatchNorm have analytically ZERO gradient: only fp32 noise (absolute tolerance)
        assert np.all(np.abs(norms - ref) <= 2e-3 * ref + 1e-5 * ref.max()), f"{tag} gradient norms differ from the reference"
        for key in g.files:
            if key.startswith(f"{tag}_grad::"):
                k = key.split("::")[1]
                assert _rel(sd[k].grad, g[key]) < 2e-3 or np.abs(g[key]).max() < 1e-5 * g[f"{tag}_grad_norms"].max(), k
            if key.startswith(f"{tag}_buf::"):
                k = key.split("::")[1]
                assert _rel(run[k], g[key]) < 1e-4, k


def test_ntxent_restatement_equals_closed_form():
    """lightly is third-party and absent: the restated CE-over-masked-logits form is pinned only by its algebraic
    definition (parity unpinned)."""
    g = torch.Generator().manual_seed(0)
    for b, D in ((4, 16), (55, 256)):
        x0, x1 = torch.randn((b, D), generator=g), torch.randn((b, D), generator=g)
        a = M.ntxent(x0.double(), x1.double()).item()
        c = M.ntxent_closed_form(x0.double(), x1.double()).item()
        assert abs(a - c) < 1e-9


def oracle_ft_run(cfg, pins=None, dtype=torch.float32):
    """Fine-tune oracle run shared with the GPU tests: mirror init -> perturbed state_dict -> oracle forward, label-smoothed
    CE (eps 0.2, ft_cls.py:145), backward."""
    model = _synth.build_ft_model(cfg)
    sd = {k: v.to(dtype) if v.dtype.is_floating_point else v for k, v in _synth.perturb_state_dict(model.state_dict(), cfg["seed"] + 10).items()}
    names = [k for k, _ in model.named_parameters()]
    for k in names:
        sd[k] = sd[k].clone().requires_grad_(True)
    for k in list(sd.keys()):
        if ".cross_attn_n." in k:
            sd[k.replace(".cross_attn_n.", ".cross_attn_1.")] = sd[k]
    pts, start, labels = _synth.ft_inputs(cfg)
    run = {}
    with M.choices(pins) as ch:
        logits = M.pc_ft_forward(sd, pts.to(dtype), start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], True, run)
        loss = M.cross_entropy_ls(logits, labels, 0.2)
        loss.backward()
    return dict(sd=sd, names=names, logits=logits.detach(), loss=loss.item(), run=run, rec=ch.rec, inputs=(pts, start, labels))


@pytest.mark.parametrize("name", ["ft_small", "ft_cfgA"])
def test_finetune_oracle_matches_reference(name, golden_dir):
    """CrossFormer_pc_mp_ft + CrossEntropyLoss(label_smoothing=0.2): oracle restatement vs vectors from the REAL reference."""
    cfg = _synth.FT_CASES[name]
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    torch.set_num_threads(8)
    o = oracle_ft_run(cfg)
    assert _rel(o["logits"], g["logits"]) < 1e-4
    assert abs(o["loss"] - float(g["loss"][0])) < 1e-4
    assert abs(F.cross_entropy(o["logits"], o["inputs"][2], label_smoothing=0.2).item() - o["loss"]) < 1e-5
    gnames = list(g["grad_names"])
    norms = {k: o["sd"][k].grad.double().norm().item() for k in gnames}
    ref = dict(zip(gnames, g["grad_norms"]))
    mx = max(ref.values())
    for k in gnames:
        assert abs(norms[k] - ref[k]) <= 2e-3 * ref[k] + 1e-5 * mx, k
    # the pre-training projection head is unused by the fine-tune forward: no gradient in the reference, zero here
    for k in o["names"]:
        if k.startswith("latent_head"):
            assert k not in gnames and (o["sd"][k].grad is None or o["sd"][k].grad.abs().max() == 0)
    for key in g.files:
        if key.startswith("grad::"):
            k = key.split("::")[1]
            assert _rel(o["sd"][k].grad, g[key]) < 2e-3 or np.abs(g[key]).max() < 1e-5 * mx, k
        if key.startswith("buf::"):
            k = key.split("::")[1]
            if k.startswith("latent_head"):      # unused head: running statistics stay at their (perturbed) initial values
                assert np.array_equal(g[key], o["sd"][k].numpy()), k
            else:
                assert _rel(o["run"][k], g[key]) < 1e-4, k


def oracle_seg_run(cfg, pins=None, drop=None, dtype=torch.float32):
    """Part-segmentation oracle run shared with the GPU tests: mirror init -> perturbed state_dict -> oracle forward,
    label-smoothed CE over all points (eps 0.2, ft_partseg.py:128,160), backward."""
    import contextlib

    model = _synth.build_seg_model(cfg)
    sd = {k: v.to(dtype) if v.dtype.is_floating_point else v for k, v in _synth.perturb_state_dict(model.state_dict(), cfg["seed"] + 10).items()}
    names = [k for k, _ in model.named_parameters()]
    for k in names:
        sd[k] = sd[k].clone().requires_grad_(True)
    for k in list(sd.keys()):
        if ".cross_attn_n." in k:
            sd[k.replace(".cross_attn_n.", ".cross_attn_1.")] = sd[k]
    pts, start, onehot, labels = _synth.seg_inputs(cfg)
    run = {}
    with M.choices(pins) as ch, (M.dropout(**drop) if drop else contextlib.nullcontext()):
        logits = M.partseg_forward(sd, pts.to(dtype), onehot.to(dtype), start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"],
                                   cfg["layer_idx"], True, run)
        loss = M.cross_entropy_ls(logits.reshape(-1, cfg["parts"]), labels.reshape(-1), 0.2)
        loss.backward()
    return dict(sd=sd, names=names, logits=logits.detach(), loss=loss.item(), run=run, rec=ch.rec,
                inputs=(pts, start, onehot, labels))


def seg_dpr_drop(cfg):
    """Injection spec of the DropPath fixture: the pinned per-sample draws the fixture generator gave the reference."""
    return dict(seed=_synth.DPR_SEED, op_bases=_synth.seg_op_bases(cfg), atten_drop=0.0, mlp_drop=0.0,
                drop_path=_synth.seg_drop_path(cfg))


@pytest.mark.parametrize("name", ["seg_small", "seg_cfgA", "seg_small_dpr"])
def test_partseg_oracle_matches_reference(name, golden_dir):
    """CrossFormer_partseg + CrossEntropyLoss(label_smoothing=0.2): oracle restatement vs vectors from the REAL reference.
    `seg_small_dpr`: max_dpr = 0.45 -- the reference's Residual applies DropPath to the WHOLE sum dropout(f(x)) + x
    (partseg.py:212); the fixture ran the reference's own Residual.forward with pinned per-sample draws."""
    cfg = _synth.SEG_CASES[name]
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    torch.set_num_threads(8)
    o = oracle_seg_run(cfg, drop=seg_dpr_drop(cfg) if cfg.get("max_dpr", 0.0) > 0 else None)
    gl = torch.from_numpy(g["logits"].astype(np.float32))
    assert _rel(o["logits"], gl) < (1e-3 if g["logits"].dtype == np.float16 else 1e-4)
    assert abs(o["loss"] - float(g["loss"][0])) < 1e-4
    gnames = list(g["grad_names"])
    assert set(gnames) == {k for k in o["names"] if o["sd"][k].grad is not None}
    if cfg.get("max_dpr", 0.0) > 0:      # the draws must matter: without them the oracle is far from the fixture
        assert _rel(oracle_seg_run(cfg)["logits"], gl) > 5e-2
    norms = {k: o["sd"][k].grad.double().norm().item() for k in gnames}
    ref = dict(zip(gnames, g["grad_norms"]))
    mx = max(ref.values())
    for k in gnames:
        assert abs(norms[k] - ref[k]) <= 2e-3 * ref[k] + 1e-5 * mx, k
    for key in g.files:
        if key.startswith("grad::"):
            k = key.split("::")[1]
            # fp32 on both sides; the weights in front of the first train-mode BatchNorm see the most cancellation (2.1e-3)
            assert _rel(o["sd"][k].grad, g[key]) < 4e-3 or np.abs(g[key]).max() < 1e-5 * mx, k
        if key.startswith("buf::"):
            k = key.split("::")[1]
            assert _rel(o["run"][k], g[key]) < 1e-4, k


@pytest.mark.parametrize("name", ["seg_small", "seg_cfgA"])
def test_partseg_mirror_parameter_names_match_reference(name, golden_dir):
    """The mirror's trainable parameters carry the reference's names (the fixture lists the real model's named_parameters
    that received a gradient), so its checkpoints load both ways (ft_partseg.py:96-104)."""
    cfg = _synth.SEG_CASES[name]
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    mine = {k for k, p in _synth.build_seg_model(cfg).named_parameters()}
    assert set(map(str, g["grad_names"])) <= mine
    # parameters of the reference without a gradient: biases nothing (all used); the mirror has no extra trainables
    assert {k for k in mine if k not in set(map(str, g["grad_names"]))} == set()
