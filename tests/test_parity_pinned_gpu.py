"""GPU: model-level parity with the DISCRETE choices pinned and with DROPOUT ON.

Round-1 gated model-level gradients at cos >= 0.90 because bf16 forward noise flips ReLU masks / max-pool arg-maxes
and each flip moves a whole gradient row.  Here the oracle (oracle/model_ref.py, fp32 CPU PyTorch, pinned to the real
reference by tests/golden/model_*.npz) is handed the choices the product's forward ACTUALLY made
(`oracle.model_ref.choices`) -- then "same choices => same gradient" must hold to bf16 GEMM accuracy:

    gate: rel-Frobenius <= 5e-2 per parameter tensor, flip rates reported separately.

Dropout-on parity (the benchmarked configuration): the product's counter-based keep-masks are restated in
oracle/rng.py and injected into the oracle (`oracle.model_ref.dropout`), attention-probability dropout
(partseg.py:81) and both Residual dropouts (partseg.py:208-213).  Gate: backbone features <= 2e-2, loss <= 5e-2,
gradients <= 5e-2 (8e-2 for tensors behind p = 0.5 masks at the small fixture).
"""
import os

import numpy as np
import pytest
import torch

import _synth
from test_oracle_model_golden import oracle_run

pytestmark = pytest.mark.gpu


def relfro(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def pins_from_tap(tap, cfg):
    """Discrete choices of one (pc forward, img forward) pair, as oracle pins.  The taps are the saved contexts of the
    product's Group2Emb / input adapter / pool+head forwards (bf16 post-ReLU activations, uint8 / int32 arg-maxes)."""
    g, a = tap["g2e"][0], tap["adapter"][0]
    heads = sorted(tap["head"], key=lambda h: -h.am.shape[0])     # pc pools 2b clouds, img pools b images (either order)
    hp, hi = heads[0], heads[1]
    c = lambda t: t.detach().cpu()
    pins = {
        "pc.g2e.relu1": c(g.h1 > 0), "pc.g2e.max2": c(g.am2).long(), "pc.g2e.relu3": c(g.h3 > 0),
        "pc.g2e.max4": c(g.am4).long(), "pc.adapter.relu": c(a.h > 0),
        "pc.pool.max": c(hp.am).long(), "pc.head.relu1": c(hp.a1 > 0), "pc.head.relu2": c(hp.a2 > 0),
        "img.pool.max": c(hi.am).long(), "img.head.relu1": c(hi.a1 > 0), "img.head.relu2": c(hi.a2 > 0),
    }
    return pins


def flip_rates(pins, rec):
    return {k: float((pins[k].reshape(-1).long() != rec[k].reshape(-1).long()).float().mean()) for k in pins}


def op_bases(pc, im):
    out = {}
    for tag, m in (("pc", pc), ("img", im)):
        out[f"{tag}.encoder.cross_attn_1"] = m.encoder.cross_attn_1._op_base
        for i, l in enumerate(m.encoder.sa_layers):
            out[f"{tag}.encoder.sa_layers.{i}"] = l._op_base
    return out


def run_product(cfg, atten_drop=0.0, mlp_drop=0.0, seed=None, linear=False):
    """Product forward + loss + backward with the choice tap on.  Returns everything the oracle needs to replay it."""
    import vipformer_b200.runtime as rt
    from vipformer_b200.loss import pretrain_loss

    o0 = oracle_run(cfg)            # weights + inputs (the unpinned run also gives the oracle's own choices)
    pc, im = _synth.build_models(cfg, atten_drop=atten_drop, mlp_drop=mlp_drop)
    pc.load_state_dict({k: v.detach() for k, v in o0["sd_pc"].items() if k in pc.state_dict()})
    im.load_state_dict({k: v.detach() for k, v in o0["sd_im"].items() if k in im.state_dict()})
    pc, im = pc.cuda().train(), im.cuda().train()
    pts, start, imgs = o0["inputs"]
    pc.fps_start_idx = torch.from_numpy(start).cuda()
    if seed is not None:
        rt.manual_seed(seed)
    rt.TAP = {}
    e0 = rt._EPOCH[0]
    try:
        pc_feats, pc_back = pc(pts.cuda())
        im_feats, im_back = im(imgs.cuda())
        tap = rt.TAP
    finally:
        rt.TAP = None
    losses = pretrain_loss(pc_feats, im_feats, temperature=0.1, cmid_weight=1.0)
    if linear:
        from test_oracle_model_golden import linear_upstream
        G = [g.cuda() for g in linear_upstream(cfg)]
        sum((t * g).sum() for t, g in zip((pc_feats, pc_back, im_feats, im_back), G)).backward()
    else:
        losses[0].backward()
    torch.cuda.synchronize()
    ob = op_bases(pc, im)
    # per-forward dropout epoch (runtime.next_op_offset): the pc encoder ran first, the img encoder second
    for k in ob:
        ob[k] += (e0 + (1 if k.startswith("pc.") else 2)) * rt.EPOCH_STRIDE
    return dict(pc=pc, im=im, tap=tap, pc_feats=pc_feats, pc_back=pc_back, im_feats=im_feats, im_back=im_back,
                losses=losses.detach().cpu().numpy(), op_bases=ob, o0=o0)


def compare_grads(r, o, tol, tol_over=None):
    bad, worst = [], 0.0
    compare_grads.errs = []
    for tag, model, sd, names in (("pc", r["pc"], o["sd_pc"], o["pnames"]), ("img", r["im"], o["sd_im"], o["inames"])):
        gmax = max(sd[k].grad.norm().item() for k in names)
        for k, p in model.named_parameters():
            ref = sd[k].grad
            if ref.norm().item() < 1e-4 * gmax:      # analytically-zero gradients (bias in front of a train-mode BN)
                if p.grad.float().norm().item() > 1e-2 * gmax:
                    bad.append((tag, k, "nonzero", p.grad.norm().item()))
                continue
            e = relfro(p.grad, ref)
            t = tol
            if tol_over:
                for frag, tv in tol_over.items():
                    if frag in k:
                        t = tv
            worst = max(worst, e)
            compare_grads.errs.append(e)
            if e > t:
                bad.append((tag, k, round(e, 4)))
    return bad, worst


PINNED_CASES = dict(_synth.MODEL_CASES, cfgA_b16=dict(_synth.MODEL_CASES["cfgA"], b=16, seed=76))


@pytest.mark.parametrize("name", ["small", "cfgA_b16"])
@pytest.mark.parametrize("loss", ["linear", "ntxent"])
def test_gradients_match_oracle_with_pinned_choices(name, loss):
    """Same discrete choices => same gradient, to bf16 accuracy.  Measured on B200 (tools/pinned_grad_sweep.py):
    loss = "linear" (a fixed linear functional of the four model outputs: every backward kernel, no loss
    amplification): worst tensor 3.7e-2 (small), 7.6e-2 (cfgA, 16 pairs; 99 % of the tensors <= 5e-2) -- the tail are the
    q/k projections (dS = P o (dP - delta) cancels) and whatever sits below the two train-mode BatchNorms of the latent
    head, which normalise over only 2b clouds / b images and amplify the forward's bf16 noise 10x (backbone features agree
    to 3e-3, projected features to 3e-2).  Gate: every tensor <= 8e-2 and >= 90 % of the tensors <= 5e-2.
    loss = "ntxent" (the real objective): NT-Xent at T = 0.1 turns that 3e-2 feature noise into 3e-1 logit noise, so the
    upstream gradient itself is only good to ~1e-1.  Gate: every tensor <= 1.5e-1 (round 1 without pins: cos >= 0.90,
    i.e. ~4.5e-1).  Flip rates are printed, not gated.
    The 4-pair golden fixture `cfgA` is not used here: its image branch normalises over FOUR samples in both latent-head
    BatchNorms, where the same comparison measures 7.5e-2 (linear, 84 % <= 5e-2) / 1.2e-1 (NT-Xent) and moves by a
    factor of two with any change of summation order -- the reference trains with 55-64 pairs per rank."""
    cfg = PINNED_CASES[name]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    r = run_product(cfg, linear=loss == "linear")
    pins = pins_from_tap(r["tap"], cfg)
    o = oracle_run(cfg, pins=pins, linear=loss == "linear")
    fr = flip_rates(pins, r["o0"]["rec"])
    print(f"[{name}] flip rates (product vs unpinned oracle):", {k: round(v, 5) for k, v in fr.items()})
    assert max(fr.values()) < 0.35            # sanity: the pins are the same kind of object as the oracle's choices
    assert relfro(r["pc_back"], o["pc_back"]) < 2e-2 and relfro(r["im_back"], o["im_back"]) < 2e-2
    assert relfro(r["pc_feats"], o["pc_feats"]) < 5e-2 and relfro(r["im_feats"], o["im_feats"]) < 5e-2
    assert np.all(np.abs(r["losses"] - np.array(o["loss"])) <= 5e-2)
    bad, worst = compare_grads(r, o, 8e-2 if loss == "linear" else 1.5e-1)
    errs = np.array(compare_grads.errs)
    print(f"[{name}/{loss}] pinned choices: worst per-parameter rel-Frobenius gradient error {worst:.4f}, "
          f"median {np.median(errs):.4f}, {100 * np.mean(errs <= 5e-2):.0f} % of tensors <= 5e-2")
    assert not bad, bad
    if loss == "linear":
        assert np.mean(errs <= 5e-2) >= 0.90, np.sort(errs)[-10:]


@pytest.mark.parametrize("name", ["small", "cfgA_b16"])
def test_dropout_on_forward_backward_match_oracle_with_injected_masks(name):
    import vipformer_b200.runtime as rt

    cfg = PINNED_CASES[name]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    seed = 0x5EED0000 + cfg["seed"]
    r = run_product(cfg, atten_drop=0.1, mlp_drop=0.5, seed=seed)
    dev_seed = int(rt.StepState.get(torch.device("cuda", torch.cuda.current_device()))[1].item())
    pins = pins_from_tap(r["tap"], cfg)
    drop = dict(seed=dev_seed, op_bases=r["op_bases"], atten_drop=0.1, mlp_drop=0.5)
    o = oracle_run(cfg, pins=pins, drop=drop)
    o_nodrop = r["o0"]
    # the injected masks matter: without them the oracle is far away
    assert relfro(r["pc_back"], o_nodrop["pc_back"]) > 0.1
    assert relfro(r["pc_back"], o["pc_back"]) < 2e-2, relfro(r["pc_back"], o["pc_back"])
    assert relfro(r["im_back"], o["im_back"]) < 2e-2, relfro(r["im_back"], o["im_back"])
    assert relfro(r["pc_feats"], o["pc_feats"]) < 5e-2 and relfro(r["im_feats"], o["im_feats"]) < 5e-2
    assert np.all(np.abs(r["losses"] - np.array(o["loss"])) <= 5e-2), (r["losses"], o["loss"])
    bad, worst = compare_grads(r, o, 1.5e-1)     # real objective: see the NT-Xent note above
    print(f"[{name}] dropout on: worst per-parameter rel-Frobenius gradient error: {worst:.4f}")
    assert not bad, bad


def test_consecutive_training_forwards_draw_different_masks():
    """Drop-in path (no engine): nn.Dropout semantics -- two consecutive training forwards differ, and the backward of
    each regenerates its own forward's masks (the seed used to stay fixed outside PretrainEngine)."""
    cfg = _synth.MODEL_CASES["small"]
    pc, _ = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc = pc.cuda().train()
    pts, start, _ = _synth.model_inputs(cfg)
    pc.fps_start_idx = torch.from_numpy(start).cuda()
    a, _ = pc(pts.cuda())
    b, _ = pc(pts.cuda())
    assert relfro(a, b) > 0.2
    pc.eval()
    with torch.no_grad():
        c, _ = pc(pts.cuda())
        d, _ = pc(pts.cuda())
    assert torch.equal(c, d)


def test_eval_mode_features_match_oracle():
    """Feature-extraction path of pretrain.py:228-276: `model(data)[1]` in eval mode (running-statistics BatchNorm,
    dropout off) at N = 1024 points -- backbone and projected features against the oracle's eval mode."""
    from oracle import model_ref as M

    cfg = dict(_synth.MODEL_CASES["cfgA"], N=1024, b=8, seed=41)
    o0 = oracle_run(dict(cfg, b=2))            # only for the perturbed weights (running stats are perturbed too)
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc.load_state_dict({k: v.detach() for k, v in o0["sd_pc"].items() if k in pc.state_dict()})
    im.load_state_dict({k: v.detach() for k, v in o0["sd_im"].items() if k in im.state_dict()})
    pc, im = pc.cuda().eval(), im.cuda().eval()
    pts, start, imgs = _synth.model_inputs(cfg)
    pc.fps_start_idx = torch.from_numpy(start).cuda()
    with torch.no_grad():
        f, bb = pc(pts.cuda())
        fi, bi = im(imgs.cuda())
        sd_pc = {k: v.detach() for k, v in o0["sd_pc"].items()}
        sd_im = {k: v.detach() for k, v in o0["sd_im"].items()}
        rf, rb = M.pc_forward(sd_pc, pts, start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], training=False)
        rfi, rbi = M.img_forward(sd_im, imgs, cfg["patch"], cfg["H"], cfg["n_sa"], training=False)
    assert bb.shape == (2 * cfg["b"], 2 * cfg["D"])
    assert relfro(bb, rb) < 2e-2 and relfro(bi, rbi) < 2e-2, (relfro(bb, rb), relfro(bi, rbi))
    assert relfro(f, rf) < 2e-2 and relfro(fi, rfi) < 2e-2, (relfro(f, rf), relfro(fi, rfi))
    # running statistics untouched by eval-mode forwards
    for k, v in pc.state_dict().items():
        if "running" in k:
            assert torch.equal(v.cpu(), o0["sd_pc"][k].detach()), k


def test_config_B_full_depth_matches_oracle():
    """E1CL8SL-H6D384-L128-MR4 at FULL depth (8 self-attention layers, 2048 points), BASELINE configs[2]: forward,
    loss, and the gradient of the fixed linear functional with pinned choices."""
    cfg = dict(D=384, H=6, n_sa=8, G=128, S=32, N=2048, MR=4, b=8, img=144, patch=12, seed=51)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    r = run_product(cfg, linear=True)
    pins = pins_from_tap(r["tap"], cfg)
    o = oracle_run(cfg, pins=pins, linear=True)
    assert r["pc_feats"].shape == (16, 384)
    assert relfro(r["pc_back"], o["pc_back"]) < 2e-2 and relfro(r["im_back"], o["im_back"]) < 2e-2
    assert np.all(np.abs(r["losses"] - np.array(o["loss"])) <= 5e-2)
    bad, worst = compare_grads(r, o, 8e-2)      # same gate as the linear case above (measured 6.3e-2)
    print(f"[cfgB] worst per-parameter rel-Frobenius gradient error with pinned choices: {worst:.4f}")
    assert not bad, bad


def test_feature_extractor_matches_eval_forward_and_oracle():
    """vipformer_b200.inference.FeatureExtractor (CUDA-graph replay, BN folded, device-resident result, ragged tail)
    == model(data)[1] in eval mode == the oracle's eval-mode backbone features (pretrain.py:228-276)."""
    from oracle import model_ref as M
    from vipformer_b200.inference import FeatureExtractor

    cfg = dict(_synth.MODEL_CASES["cfgA"], N=1024, b=2, seed=43)
    o0 = oracle_run(cfg)
    pc, _ = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc.load_state_dict({k: v.detach() for k, v in o0["sd_pc"].items() if k in pc.state_dict()})
    n, B = 21, 8
    pts = torch.from_numpy(_synth.make_clouds("randn", n, 1024, 77))
    start = torch.from_numpy(_synth.make_start(B, 1024, 77))
    outs = []
    for graph in (True, False):
        fx = FeatureExtractor(pc, B, 1024, use_cuda_graph=graph)
        fx.fixed_start = start.cuda()
        outs.append(fx(pts.pin_memory()))
    assert outs[0].shape == (n, 2 * cfg["D"]) and torch.equal(outs[0], outs[1])
    sd = {k: v.detach() for k, v in o0["sd_pc"].items()}
    with torch.no_grad():
        for i0 in range(0, n, B):
            m = min(B, n - i0)
            _, rb = M.pc_forward(sd, pts[i0:i0 + m], start[:m].numpy(), cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], training=False)
            assert relfro(outs[0][i0:i0 + m], rb) < 2e-2, (i0, relfro(outs[0][i0:i0 + m], rb))
    f, l = FeatureExtractor(pc, B, 1024).extract([(pts[:8], torch.arange(8)), (pts[8:13], torch.arange(5))])
    assert f.shape == (13, 2 * cfg["D"]) and l.tolist() == list(range(8)) + list(range(5)) and np.isfinite(f).all()
