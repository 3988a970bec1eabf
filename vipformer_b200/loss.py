"""NT-Xent contrastive objective: drop-in for `lightly.loss.NTXentLoss(temperature)` as used at pretrain.py:155,196,202,
plus the fused loss composition of pretrain.py:189-207.

Kernels (csrc/loss_optim.cu): row normalisation, fp32 similarity GEMM into a small logits scratch ([2b, 2bW]; 1 MB
at b = 256), masked row log-sum-exp + reduction; backward = weights in place, fp32 GEMM, normalisation backward.  With gather_distributed=True the columns are the
all-gathered embeddings of every rank (NCCL all-gather of the L2-normalised rows; the backward needs only an
all-gather of the per-row log-sum-exp vector) so negatives span the global batch.
"""
import torch
import torch.nn as nn

from . import _lib, ops

F32 = torch.float32


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_layout(rank, world, b):
    """Column layout of the all-gathered embeddings: cat(all ranks' out0 blocks, all ranks' out1 blocks).
    Returns (col_offset, half): local row i < b sits at column col_offset + i, row b + i at half + col_offset + i."""
    return rank * b, world * b


def self_pos_columns(i, b, col_offset, half):
    """(own column, positive column) of local row i -- the host-side statement of `self_pos` in loss_optim.cu."""
    if i < b:
        return col_offset + i, half + col_offset + i
    return half + col_offset + (i - b), col_offset + (i - b)


def _stack2(a, b):
    """[a; b] as one contiguous fp32 [2b, D] buffer (copy kernels, no ATen cat)."""
    n, D = a.shape
    x = torch.empty((2 * n, D), dtype=F32, device=a.device)
    ops.add_scale(a.float().contiguous(), None, 1.0, out=x[:n])
    ops.add_scale(b.float().contiguous(), None, 1.0, out=x[n:])
    return x


class _NTXentCore:
    """forward/backward on a stacked [2b, D] input; shared by the autograd functions below."""

    @staticmethod
    def fwd(x, temperature, gather):
        n, D = x.shape
        b = n // 2
        z, norm = ops.l2norm_rows(x)
        dist = _dist() if gather else None
        if dist is not None:
            W, r = dist.get_world_size(), dist.get_rank()
            zc = torch.empty((2 * W * b, D), dtype=F32, device=x.device)
            dist.all_gather_into_tensor(zc[:W * b], z[:b].contiguous())
            dist.all_gather_into_tensor(zc[W * b:], z[b:].contiguous())
            col_offset, half = shard_layout(r, W, b)
        else:
            zc, col_offset, half = z, 0, b
        loss = ops.zeros_(torch.empty(1, dtype=F32, device=x.device))
        lse, S = ops.ntxent_fwd(z, zc, b, col_offset, half, temperature, loss)
        return loss, [z, norm, zc, lse, b, col_offset, half, temperature, dist, S]

    @staticmethod
    def bwd(saved, gscale, upstream=None):
        z, norm, zc, lse, b, col_offset, half, temperature, dist, S = saved
        if S is None:
            raise RuntimeError("NT-Xent backward ran twice on one forward: its logits scratch is overwritten in place by "
                               "the first backward (retain_graph=True is not supported by this fused loss)")
        if dist is not None:
            lse_all = torch.empty(2 * half, dtype=F32, device=z.device)
            dist.all_gather_into_tensor(lse_all[:half], lse[:b].contiguous())
            dist.all_gather_into_tensor(lse_all[half:], lse[b:].contiguous())
        else:
            lse_all = lse
        saved[-1] = None
        # DDP averages parameter gradients over ranks, so the per-rank seed stays 1/(2b) for any world size
        return ops.ntxent_bwd(z, norm, zc, lse_all, S, b, col_offset, half, temperature, gscale / (2 * b), upstream)


class _NTXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out0, out1, temperature, gather):
        _lib.require_cuda(out0, out1)
        if out0.shape != out1.shape or out0.dim() != 2:
            raise ValueError("NTXentLoss expects two [batch, dim] tensors of equal shape")
        loss, saved = _NTXentCore.fwd(_stack2(out0, out1), temperature, gather)
        ctx.saved = saved
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        up = dloss.float().reshape(1).contiguous()
        dx = _NTXentCore.bwd(ctx.saved, 1.0, up)
        b = ctx.saved[4]
        return dx[:b], dx[b:], None, None


class NTXentLoss(nn.Module):
    """`lightly.loss.NTXentLoss`-compatible: forward(out0 [b,D], out1 [b,D]) -> 0-dim fp32 loss (with autograd)."""

    def __init__(self, temperature: float = 0.5, memory_bank_size: int = 0, gather_distributed: bool = False):
        super().__init__()
        if memory_bank_size != 0:
            raise NotImplementedError("memory bank negatives are not used by ViPFormer (pretrain.py:155) and not built")
        if abs(temperature) < 1e-8:
            raise ValueError(f"Illegal temperature: abs({temperature}) < 1e-8")
        self.temperature = temperature
        self.gather_distributed = gather_distributed

    def forward(self, out0, out1):
        return _NTXentFn.apply(out0, out1, float(self.temperature), bool(self.gather_distributed))


class _PretrainLossFn(torch.autograd.Function):
    """total = NTXent(t1, t2) + w * NTXent((t1+t2)/2, img)   (pretrain.py:189-207, modality 'both')."""

    @staticmethod
    def forward(ctx, pc_feats, img_feats, temperature, cmid_weight, gather):
        _lib.require_cuda(pc_feats, img_feats)
        pc = pc_feats.float().contiguous()
        b = pc.shape[0] // 2
        l_imid, s_imid = _NTXentCore.fwd(pc, temperature, gather)          # rows already stacked as [t1; t2]
        avg = ops.add_scale(pc[:b], pc[b:], 0.5)
        l_cmid, s_cmid = _NTXentCore.fwd(_stack2(avg, img_feats), temperature, gather)
        total = torch.empty(3, dtype=F32, device=pc.device)
        ops.add_scale(l_imid, None, 1.0, out=total[1:2])
        ops.add_scale(l_cmid, None, 1.0, out=total[2:3])
        ops.add_scale(l_imid, ops.add_scale(l_cmid, None, float(cmid_weight)), 1.0, out=total[0:1])
        ctx.saved = (s_imid, s_cmid, b, float(cmid_weight))
        return total

    @staticmethod
    def backward(ctx, dtotal):
        s_imid, s_cmid, b, w = ctx.saved
        up = dtotal.float().contiguous()[0:1]      # gradients flow through total[0] only (entries 1,2 are for logging)
        d_imid = _NTXentCore.bwd(s_imid, 1.0, up)                 # [2b, D] w.r.t. [t1; t2]
        d_cmid = _NTXentCore.bwd(s_cmid, w, up)                   # [2b, D] w.r.t. [(t1+t2)/2; img]
        half = ops.add_scale(d_cmid[:b], None, 0.5)
        dpc = torch.empty_like(d_imid)
        ops.add_scale(d_imid[:b], half, 1.0, out=dpc[:b])
        ops.add_scale(d_imid[b:], half, 1.0, out=dpc[b:])
        return dpc, d_cmid[b:], None, None, None


def pretrain_loss(pc_feats, img_feats, temperature=0.1, cmid_weight=1.0, gather_distributed=False):
    """-> fp32 tensor [3] = (total, loss_imid, loss_cmid); backpropagate through element 0."""
    return _PretrainLossFn.apply(pc_feats, img_feats, float(temperature), float(cmid_weight), bool(gather_distributed))
