"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path.

The kernels need a GPU, so what runs here is the HOST statement of what `loss.py` / `loss_optim.cu` do across ranks --
the shard layout of the all-gathered embeddings, the per-row (self, positive) columns, the symmetric gradient formula
that needs only an all-gather of the per-row log-sum-exp, and the 1/(2b) seed under a mean all-reduce -- checked against
the single-process oracle loss and its autograd on the concatenated batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model_ref as M
from vipformer_b200.loss import self_pos_columns, shard_layout

B, D, T = 6, 32, 0.1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x0_all, x1_all = torch.randn((world * B, D), generator=g, dtype=torch.float64), torch.randn((world * B, D), generator=g, dtype=torch.float64)
    x0, x1 = x0_all[rank * B:(rank + 1) * B], x1_all[rank * B:(rank + 1) * B]
    x = torch.cat([x0, x1], 0)
    norm = x.norm(dim=1).clamp_min(1e-12)
    z = x / norm[:, None]
    # all-gather exactly as _NTXentCore.fwd lays it out
    col_offset, half = shard_layout(rank, world, B)
    n_c = 2 * B * world
    parts = [torch.empty((2 * B, D), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, z.contiguous())          # ONE all-gather of the stacked [out0; out1] rows: rank-major columns
    zc = torch.cat(parts)
    s = z @ zc.t() / T
    lse = torch.empty(2 * B, dtype=torch.float64)
    loss = 0.0
    for i in range(2 * B):
        me, pos = self_pos_columns(i, B, col_offset, half)
        mask = torch.ones(n_c, dtype=torch.bool)
        mask[me] = False
        lse[i] = torch.logsumexp(s[i][mask], 0)
        loss += (lse[i] - s[i, pos]) / (2 * B)
    # backward: all-gather of the per-row LSE, then each rank alone computes d(sum over ALL rows)/d z_k for its rows
    ls = [torch.empty(2 * B, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(ls, lse.contiguous())           # ONE all-gather, same rank-major layout as the columns
    lse_all = torch.cat(ls)
    gscale = 1.0 / (2 * B)          # stays 1/(2b): the gradient all-reduce AVERAGES over ranks
    dx = torch.empty_like(x)
    for k in range(2 * B):
        me, pos = self_pos_columns(k, B, col_offset, half)
        w = torch.exp(s[k] - lse[k]) + torch.exp(s[k] - lse_all)
        w[me] = 0.0
        w[pos] -= 2.0
        gk = gscale / T * (w[:, None] * zc).sum(0)
        dx[k] = (gk - z[k] * (z[k] @ gk)) / norm[k]
    # "DDP": parameters are the inputs themselves here; mean all-reduce of a per-rank scatter of dx
    full = torch.zeros((2, world * B, D), dtype=torch.float64)
    full[0, rank * B:(rank + 1) * B], full[1, rank * B:(rank + 1) * B] = dx[:B], dx[B:]
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    full /= world
    losses = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(losses, torch.tensor([float(loss)], dtype=torch.float64))
    if rank == 0:
        ret["loss"] = float(torch.stack(losses).mean())
        ret["g0"], ret["g1"] = full[0].clone(), full[1].clone()
        ret["x0"], ret["x1"] = x0_all, x1_all
    dist.destroy_process_group()


def test_sharded_global_ntxent_equals_single_process():
    world = 2
    port = 29600 + os.getpid() % 300
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        x0, x1 = ret["x0"].clone().requires_grad_(True), ret["x1"].clone().requires_grad_(True)
        ref = M.ntxent(x0, x1, T)
        ref.backward()
        assert abs(ret["loss"] - ref.item()) < 1e-10
        assert torch.allclose(ret["g0"], x0.grad, atol=1e-10) and torch.allclose(ret["g1"], x1.grad, atol=1e-10)


def test_shard_layout_covers_every_column_once():
    for world in (1, 2, 4, 8):
        seen = set()
        for r in range(world):
            off, half = shard_layout(r, world, B)
            assert half == B and off == 2 * r * B
            for i in range(2 * B):
                me, pos = self_pos_columns(i, B, off, half)
                assert abs(me - pos) == half and me not in seen
                seen.add(me)
        assert seen == set(range(2 * world * B))
