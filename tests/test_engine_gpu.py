"""GPU: the pre-training engine (flat arena, fused AdamW, CUDA graph) on one GPU."""
import pytest
import torch

import _synth

pytestmark = pytest.mark.gpu


def _engine(graph, seed=3, lr=1e-3, **kw):
    from vipformer_b200.engine import PretrainEngine

    cfg = _synth.MODEL_CASES["small"]
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    eng = PretrainEngine(pc, im, batch_pairs=cfg["b"], num_points=cfg["N"], lr=lr, use_cuda_graph=graph, seed=seed, **kw)
    pts, _, imgs = _synth.model_inputs(cfg)
    eng.pc_in.copy_(pts.cuda())
    eng.img_in.copy_(imgs.cuda().permute(0, 3, 1, 2))
    return eng, cfg, pts, imgs


@pytest.mark.parametrize("graph", [False, True])
def test_engine_steps_train(graph):
    eng, cfg, pts, imgs = _engine(graph)
    p0 = eng.arena.flat_p.clone()
    hist = [eng.step().cpu().clone() for _ in range(8)]
    torch.cuda.synchronize()
    assert all(torch.isfinite(h).all() for h in hist)
    assert not torch.equal(p0, eng.arena.flat_p)
    # bf16 shadows track the fp32 masters after every fused AdamW step
    assert torch.equal(eng.arena.flat_bf.float(), eng.arena.flat_p.to(torch.bfloat16).float())
    assert eng.state[0].item() == eng.steps_done == 8     # graph warm-up does not train (one step per batch)
    # parameters are views of the arena; gradients too
    p = next(eng.pc_model.parameters())
    assert p.data_ptr() == eng.arena.flat_p.data_ptr() and p.grad.data_ptr() == eng.arena.flat_g.data_ptr()
    # total = imid + cmid_weight * cmid
    assert abs(hist[-1][0] - (hist[-1][1] + hist[-1][2])) < 1e-4
    # same batch every step: the objective must go down
    assert hist[-1][0] < hist[0][0]


def test_engine_host_step_and_state_dict_roundtrip():
    eng, cfg, pts, imgs = _engine(False)
    b = cfg["b"]
    out = eng.step_host(pts[:b].contiguous().pin_memory(), pts[b:].contiguous().pin_memory(),
                        imgs.permute(0, 3, 1, 2).contiguous().pin_memory())
    assert out.shape == (3,) and torch.isfinite(out).all()
    sd = eng.pc_model.state_dict()
    pc2, _ = _synth.build_models(cfg)
    pc2.load_state_dict(sd)          # reference-layout keys
    assert set(sd.keys()) == set(pc2.state_dict().keys())


@pytest.mark.parametrize("graph", [False, True])
def test_engine_two_stream_overlap_matches_serial(graph):
    """The image branch on a side stream (engine default) must give the serial schedule's results: same weights, same
    dropout seeds, same kernels, only the interleaving differs. lr = 0 keeps the weights fixed, so every step of the two
    runs is comparable (with lr > 0 Adam turns last-bit differences of the split-K fp32 atomics into +-lr updates and
    even two serial runs drift apart); the tolerance covers those atomics."""
    runs = []
    for overlap in (False, True):
        torch.manual_seed(0)                              # same initial weights in both runs
        eng, *_ = _engine(graph, seed=11, lr=0.0, overlap_branches=overlap)
        hist = torch.stack([eng.step().clone() for _ in range(4)]).cpu()
        torch.cuda.synchronize()
        runs.append((hist, eng.arena.flat_g.clone()))
    (h0, g0), (h1, g1) = runs
    assert torch.allclose(h0, h1, rtol=1e-5, atol=1e-5), (h0, h1)
    assert not torch.allclose(h0[0], h0[1], rtol=1e-4, atol=1e-4)   # fresh dropout masks every step
    assert ((g0 - g1).norm() / g0.norm()).item() < 5e-4   # fp32 atomics of the split-K reductions


def test_engine_prefetched_host_steps_match_synchronous():
    """step_host_prefetched (H2D of the next batch on a copy stream) must consume exactly the batches it was given, in
    order: same losses as the synchronous step_host on the same sequence of batches (lr = 0, fixed weights)."""
    cfg = _synth.MODEL_CASES["small"]
    b = cfg["b"]
    batches = []
    for k in range(3):
        g = torch.Generator().manual_seed(50 + k)
        pts = torch.randn((2 * b, cfg["N"], 3), generator=g) * 0.4
        imgs = torch.randn((b, 3, 144, 144), generator=g)
        batches.append((pts[:b].contiguous().pin_memory(), pts[b:].contiguous().pin_memory(), imgs.pin_memory()))
    runs = []
    for mode in ("sync", "prefetch"):
        torch.manual_seed(0)
        eng, *_ = _engine(False, seed=5, lr=0.0)
        out = []
        if mode == "sync":
            for bt in batches:
                out.append(eng.step_host(*bt).clone())
        else:
            eng.prefetch_host(*batches[0])
            for k in range(3):
                out.append(eng.step_host_prefetched(batches[k + 1] if k + 1 < 3 else None).clone())
        runs.append(torch.stack(out))
    # lagged loss read-back: call k returns the losses of step k-1, drain_losses() the last
    torch.manual_seed(0)
    eng, *_ = _engine(False, seed=5, lr=0.0)
    eng.prefetch_host(*batches[0])
    lagged = [eng.step_host_prefetched(batches[k + 1] if k + 1 < 3 else None, lag_losses=True) for k in range(3)]
    assert lagged[0] is None
    lagged = [t.clone() for t in lagged[1:]] + [eng.drain_losses().clone()]
    assert eng.drain_losses() is None
    runs.append(torch.stack(lagged))
    assert torch.allclose(runs[0], runs[1], rtol=1e-5, atol=1e-5), runs
    assert torch.allclose(runs[0], runs[2], rtol=1e-5, atol=1e-5), runs
    assert not torch.allclose(runs[0][0], runs[0][1], rtol=1e-3, atol=1e-3)   # the batches do differ


def test_engine_step_matches_oracle_fwd_bwd_adamw():
    """One engine step (dropout ON, eager) against the oracle: forward + both NT-Xent terms + backward with the product's
    discrete choices pinned and its dropout masks injected, then torch.optim.AdamW (pretrain.py:121-124,209-211).
    Adam's first step is p -= lr * (g / (|g| + eps) + wd * p): the update is +-lr wherever |g| >> eps, so parameters are
    compared through (i) the first moment m = 0.1 g (rel-Frobenius <= 1e-1 per tensor, measured <= 5.3e-2) and (ii) the
    sign agreement / cosine of the parameter deltas."""
    import os

    import numpy as np

    import vipformer_b200.runtime as rt
    from test_oracle_model_golden import oracle_run
    from test_parity_pinned_gpu import op_bases, pins_from_tap, relfro
    from vipformer_b200.engine import PretrainEngine

    cfg = _synth.MODEL_CASES["small"]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    o0 = oracle_run(cfg)
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
    pc.load_state_dict({k: v.detach() for k, v in o0["sd_pc"].items() if k in pc.state_dict()})
    im.load_state_dict({k: v.detach() for k, v in o0["sd_im"].items() if k in im.state_dict()})
    eng = PretrainEngine(pc, im, batch_pairs=cfg["b"], num_points=cfg["N"], lr=1e-3, use_cuda_graph=False, seed=9)
    pts, _, imgs = o0["inputs"]
    eng.pc_in.copy_(pts.cuda())
    eng.img_in.copy_(imgs.cuda().permute(0, 3, 1, 2))
    p_old = {k: v.detach().clone().cpu() for m in (("pc", eng.pc_model), ("img", eng.img_model)) for k, v in
             ((m[0] + "." + n, p) for n, p in m[1].named_parameters())}
    rt.TAP = {}
    try:
        losses = eng.step().cpu().numpy()
        tap = rt.TAP
    finally:
        rt.TAP = None
    torch.cuda.synchronize()
    seed = int(eng.state[1].item())
    pins = pins_from_tap(tap, cfg)
    drop = dict(seed=seed, op_bases=op_bases(eng.pc_model, eng.img_model), atten_drop=0.1, mlp_drop=0.5)
    o = oracle_run(cfg, pins=pins, drop=drop, start=eng.start.cpu().numpy())
    assert np.all(np.abs(losses - np.array(o["loss"])) <= 5e-2), (losses, o["loss"])
    leaves = [o["sd_pc"][k] for k in o["pnames"]] + [o["sd_im"][k] for k in o["inames"]]
    grads = {("pc." + k): o["sd_pc"][k].grad.clone() for k in o["pnames"]}
    grads.update({("img." + k): o["sd_im"][k].grad.clone() for k in o["inames"]})
    opt = torch.optim.AdamW(leaves, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    opt.step()
    new_ref = {("pc." + k): o["sd_pc"][k].detach() for k in o["pnames"]}
    new_ref.update({("img." + k): o["sd_im"][k].detach() for k in o["inames"]})
    # engine first moments, per parameter (arena order = root.parameters() order)
    names = ["pc." + n for n, _ in eng.pc_model.named_parameters()] + ["img." + n for n, _ in eng.img_model.named_parameters()]
    plist = list(eng.pc_model.parameters()) + list(eng.img_model.parameters())
    assert [id(p) for p in plist] == [id(p) for p in eng.arena.params()]
    gmax = max(g.norm().item() for g in grads.values())
    bad, agree_n, agree_d = [], 0, 0
    for n, p, off in zip(names, plist, eng.arena._offsets):
        g = grads[n]
        if g.norm().item() < 1e-4 * gmax:
            continue
        m_eng = eng.m[off:off + p.numel()].view(p.shape).cpu()
        e = relfro(m_eng, 0.1 * g)
        if e > 1e-1:         # NT-Xent upstream at T = 0.1, 6 pairs: see tests/test_parity_pinned_gpu.py (measured <= 5.3e-2)
            bad.append((n, round(e, 4)))
        d_eng = (p.detach().cpu() - p_old[n]).reshape(-1)
        d_ref = (new_ref[n] - p_old[n]).reshape(-1)
        big = g.reshape(-1).abs() > 0.05 * g.abs().mean()
        agree_n += int((torch.sign(d_eng[big]) == torch.sign(d_ref[big])).sum())
        agree_d += int(big.sum())
        # AdamW's first step is lr * sign(g) per element: elements whose gradient is noise-sized carry a random sign on
        # both sides, so the direction is compared on the elements with a gradient worth the name (the same `big` mask)
        if int(big.sum()) < 8:
            continue
        cos = torch.nn.functional.cosine_similarity(d_eng[big].double().view(1, -1), d_ref[big].double().view(1, -1)).item()
        if cos < 0.93:
            bad.append((n, "delta cos", round(cos, 4)))
    assert not bad, bad
    assert agree_n / agree_d > 0.985, agree_n / agree_d


def test_multi_gpu_equivalence_over_nccl():
    """N-rank == single process on the concatenated batch (loss, feature gradients), parameters stay bit-identical across
    ranks, eager == graph: tests/mp_equiv.py under torchrun, when the box has >= 2 GPUs."""
    import os
    import subprocess
    import sys

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(root, "tests", "mp_equiv.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "MP_EQUIV_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
