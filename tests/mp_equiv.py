"""torchrun script (NOT collected by pytest): N-rank hardware equivalence over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_equiv.py

(1) loss level: every rank holds its shard of projected features; the product's global-negative NT-Xent pair
    (all-gather of normalised rows + of the row LSEs, loss.py) must give  mean_r loss_r == the single-process oracle loss
    on the CONCATENATED batch, and rank r's feature gradient == W * d(global loss)/d(features of rank r)  (DDP then
    averages parameter gradients over ranks, pretrain.py:104-105).
(2) engine level: ranks train on different batches; after the flat gradient all-reduce + AdamW the parameters must be
    bit-identical on every rank, the all-reduced flat gradient must equal the mean of the ranks' local gradients, and
    eager and CUDA-graph steps must agree.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _say(rank, msg):
    if os.environ.get("VPF_MP_VERBOSE"):
        print(f"[mp_equiv rank {rank}] {msg}", file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import _synth
    from oracle import model_ref as M
    from vipformer_b200.engine import PretrainEngine
    from vipformer_b200.loss import pretrain_loss

    # ---------------------------------------------------------------- (1) loss level
    b, D = 24, 256
    g = torch.Generator().manual_seed(7)
    t1, t2, im = (torch.randn((world * b, D), generator=g, dtype=torch.float64) for _ in range(3))
    sl = slice(rank * b, (rank + 1) * b)
    pc_local = torch.cat([t1[sl], t2[sl]]).float().cuda().requires_grad_(True)
    im_local = im[sl].float().cuda().requires_grad_(True)
    losses = pretrain_loss(pc_local, im_local, 0.1, 1.0, gather_distributed=True)
    losses[0].backward()
    mean_loss = losses.detach().clone()
    dist.all_reduce(mean_loss, op=dist.ReduceOp.AVG)
    t1r, t2r, imr = (t.clone().requires_grad_(True) for t in (t1, t2, im))
    ref_total, ref_imid, ref_cmid = M.pretrain_loss(torch.cat([t1r, t2r]), imr)
    ref_total.backward()
    ref = np.array([ref_total.item(), ref_imid.item(), ref_cmid.item()])
    assert np.allclose(mean_loss.cpu().numpy(), ref, rtol=2e-5, atol=2e-5), (mean_loss, ref)
    gref_pc = world * torch.cat([t1r.grad[sl], t2r.grad[sl]])
    gref_im = world * imr.grad[sl]
    rel = lambda a, r: ((a.detach().double().cpu() - r).norm() / r.norm()).item()
    assert rel(pc_local.grad, gref_pc) < 1e-4 and rel(im_local.grad, gref_im) < 1e-4, (rel(pc_local.grad, gref_pc), rel(im_local.grad, gref_im))

    _say(rank, "loss level ok")
    # ---------------------------------------------------------------- (2) engine level
    cfg = _synth.MODEL_CASES["small"]
    results = {}
    for graph in (False, True):
        torch.manual_seed(100 + rank)           # deliberately DIFFERENT initial weights per rank: the engine broadcasts rank 0's
        pc, img = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
        eng = PretrainEngine(pc, img, batch_pairs=cfg["b"], num_points=cfg["N"], lr=1e-3, use_cuda_graph=graph, seed=3)
        gg = torch.Generator().manual_seed(50 + rank)          # different data per rank
        eng.pc_in.copy_((torch.randn((2 * cfg["b"], cfg["N"], 3), generator=gg) * 0.4).cuda())
        eng.img_in.copy_(torch.randn((cfg["b"], 3, 144, 144), generator=gg).cuda())
        hist = []
        for i in range(3):
            hist.append(eng.step().clone())
            torch.cuda.synchronize()
            _say(rank, f"graph={graph} step {i} done")
        assert all(torch.isfinite(h).all() for h in hist)
        p = eng.arena.flat_p
        p0 = p.clone()
        dist.broadcast(p0, src=0)
        assert torch.equal(p, p0), f"parameters diverged across ranks (graph={graph})"
        gsum = eng.arena.flat_g.clone()
        g0 = gsum.clone()
        dist.broadcast(g0, src=0)
        assert torch.equal(gsum, g0), "all-reduced gradients differ across ranks"
        assert eng.state[0].item() == 3
        results[graph] = (torch.stack(hist).cpu(), p.clone())
        eng.close()
    # eager vs graph: same seeds, same data; the only differences are fp32 atomics order
    h0, h1 = results[False][0], results[True][0]
    assert torch.allclose(h0[0], h1[0], rtol=1e-3, atol=1e-3), (h0, h1)
    dist.barrier()
    if rank == 0:
        print(f"MP_EQUIV_OK world={world} loss={mean_loss.cpu().tolist()} engine_losses={h0[-1].tolist()}", flush=True)
    # tear-down: the CUDA graph holds captured NCCL kernels -- drop it (and the engines) before the communicator
    torch.cuda.synchronize()
    del eng, results, pc, img
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
