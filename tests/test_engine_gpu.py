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
