"""NT-Xent contrastive objective: drop-in for `lightly.loss.NTXentLoss(temperature)` as used at pretrain.py:155,196,202,
plus the fused loss composition of pretrain.py:189-207.

Kernels (csrc/loss_optim.cu): row normalisation, fp32 similarity GEMM into a small logits scratch ([2b, 2bW]; 1 MB
at b = 256), masked row log-sum-exp + reduction; backward = weights in place, fp32 GEMM, normalisation backward.  With
gather_distributed=True the columns are the all-gathered embeddings of every rank so negatives span the global batch:
per step ONE NCCL all-gather of the packed L2-normalised rows of both loss terms and ONE of their per-row log-sum-exps
(SURVEY.md 8e), both issued in the forward pass -- the backward is collective-free.
"""
import torch
import torch.nn as nn

from . import _lib, ops

F32 = torch.float32
_SERIAL_TERMS = __import__("os").environ.get("VPF_NTXENT_SERIAL", "0") == "1"


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_layout(rank, world, b):
    """Column layout of the all-gathered embeddings (rank-major): rank r contributes the 2b columns
    [2rb, 2rb + b) = its out0 rows and [2rb + b, 2rb + 2b) = its out1 rows.
    Returns (col_offset, half): local row i < b sits at column col_offset + i, row b + i at col_offset + half + i."""
    return 2 * rank * b, b


def self_pos_columns(i, b, col_offset, half):
    """(own column, positive column) of local row i -- the host-side statement of `self_pos` in loss_optim.cu."""
    if i < b:
        return col_offset + i, half + col_offset + i
    return half + col_offset + (i - b), col_offset + (i - b)


def _copy_rows(src, dst):
    ops.add_scale(src.float().contiguous(), None, 1.0, out=dst)


class _NTXentPack:
    """forward/backward of `nseg` NT-Xent terms that share ONE all-gather of the normalised embeddings and ONE of the
    per-row log-sum-exps (both issued in the FORWARD pass, on the caller's thread and stream: nothing collective runs
    inside autograd's backward, which is what lets the whole step be captured in a CUDA graph with world_size > 1).

    x_pack [nseg * 2b, D]: segment s holds the stacked [out0; out1] rows of term s.  Every rank gathers the packed
    block, so logits column j of term s lives at row (j // 2b) * (nseg * 2b) + s * 2b + j % 2b of the gathered buffer
    (the ColMap of loss_optim.cu) -- no re-layout copy."""

    @staticmethod
    def fwd(x_pack, nseg, temperature, gather):
        n, D = x_pack.shape
        n_r = n // nseg
        b = n_r // 2
        dev = x_pack.device
        z, norm = ops.l2norm_rows(x_pack)
        dist = _dist() if gather else None
        W, r = (dist.get_world_size(), dist.get_rank()) if dist is not None else (1, 0)
        if dist is not None:
            zc = torch.empty((W * n, D), dtype=F32, device=dev)
            dist.all_gather_into_tensor(zc, z)
        else:
            zc = z
        col_offset, half = shard_layout(r, W, b)
        n_c = W * n_r
        # every term in one launch per stage: term s reads columns through the map (blk = 2b, ld = nseg * 2b, base = s * 2b)
        losses, lse, Sall = ops.ntxent_pack_fwd(z, nseg, zc, b, col_offset, half, temperature, n_c, (n_r, n, 0), n_r)
        S = [Sall[s] for s in range(nseg)]
        if _SERIAL_TERMS:       # measurement switch (tools/ablate_step.py): one launch chain per term, as before the packing
            for s in range(1, nseg):
                ops.ntxent_fwd(z[s * n_r:(s + 1) * n_r], zc, b, col_offset, half, temperature, losses[s:s + 1],
                               n_c=n_c, colmap=(n_r, n, s * n_r), lse_out=lse[s * n_r:(s + 1) * n_r])
        if dist is not None:
            lse_all = torch.empty(W * n, dtype=F32, device=dev)
            dist.all_gather_into_tensor(lse_all, lse)
        else:
            lse_all = lse
        return losses, dict(z=z, norm=norm, zc=zc, lse_all=lse_all, S=S, Sall=Sall, nseg=nseg, b=b, n_r=n_r, n=n, n_c=n_c, col_offset=col_offset,
                            half=half, T=temperature)

    @staticmethod
    def bwd(sv, seg, gscale, upstream=None):
        """-> gradient [2b, D] w.r.t. segment `seg`'s stacked rows."""
        S = sv["S"][seg]
        if S is None:
            raise RuntimeError("NT-Xent backward ran twice on one forward: its logits scratch is overwritten in place by "
                               "the first backward (retain_graph=True is not supported by this fused loss)")
        sv["S"][seg] = None
        n_r, b = sv["n_r"], sv["b"]
        sl = slice(seg * n_r, (seg + 1) * n_r)
        # DDP averages parameter gradients over ranks, so the per-rank seed stays 1/(2b) for any world size
        return ops.ntxent_bwd(sv["z"][sl], sv["norm"][sl], sv["zc"], sv["lse_all"], S, b, sv["col_offset"], sv["half"],
                              sv["T"], gscale / (2 * b), upstream, n_c=sv["n_c"], colmap=(n_r, sv["n"], seg * n_r))


    @staticmethod
    def bwd_all(sv, gscales, upstream=None):
        """-> gradient [nseg * 2b, D] w.r.t. the packed rows, every term in one launch per stage."""
        if any(S is None for S in sv["S"]):
            raise RuntimeError("NT-Xent backward ran twice on one forward: its logits scratch is overwritten in place by "
                               "the first backward (retain_graph=True is not supported by this fused loss)")
        nseg, n_r, b = sv["nseg"], sv["n_r"], sv["b"]
        sv["S"] = [None] * nseg
        return ops.ntxent_pack_bwd(sv["z"], sv["norm"], nseg, sv["zc"], sv["lse_all"], sv["Sall"], b, sv["col_offset"],
                                   sv["half"], sv["T"], [g / (2 * b) for g in gscales], upstream, sv["n_c"],
                                   (n_r, sv["n"], 0), n_r)


class _NTXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out0, out1, temperature, gather):
        _lib.require_cuda(out0, out1)
        if out0.shape != out1.shape or out0.dim() != 2:
            raise ValueError("NTXentLoss expects two [batch, dim] tensors of equal shape")
        b, D = out0.shape
        x = torch.empty((2 * b, D), dtype=F32, device=out0.device)
        _copy_rows(out0, x[:b])
        _copy_rows(out1, x[b:])
        loss, saved = _NTXentPack.fwd(x, 1, temperature, gather)
        ctx.saved = saved
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        up = dloss.float().reshape(1).contiguous()
        dx = _NTXentPack.bwd(ctx.saved, 0, 1.0, up)
        b = ctx.saved["b"]
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


class _CELabelSmoothingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, eps):
        _lib.require_cuda(logits, labels)
        if logits.dim() != 2:
            raise ValueError("CrossEntropyLoss expects [batch, classes] logits")
        x = logits if (logits.dtype == F32 and logits.stride(1) == 1) else logits.float().contiguous()
        loss, d = ops.ce_label_smoothing(x, labels.reshape(-1).long(), eps)
        ctx.d = d
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        d = ctx.d
        up = dloss.float().reshape(1)
        out = torch.empty_like(d)
        _lib.call("vpf_scale_by", ops._p(d), ops._p(up), ops._p(out), ops._ll(d.numel()), ops._s())
        return out, None, None


class CrossEntropyLoss(nn.Module):
    """`torch.nn.CrossEntropyLoss(label_smoothing=eps)` (mean reduction) as used by ft_cls.py:145,176."""

    def __init__(self, label_smoothing: float = 0.0):
        super().__init__()
        if not 0.0 <= label_smoothing < 1.0:
            raise ValueError("label_smoothing must be in [0, 1)")
        self.label_smoothing = float(label_smoothing)

    def forward(self, input, target):
        return _CELabelSmoothingFn.apply(input, target, self.label_smoothing)


class _PretrainLossFn(torch.autograd.Function):
    """total = NTXent(t1, t2) + w * NTXent((t1+t2)/2, img)   (pretrain.py:189-207, modality 'both')."""

    @staticmethod
    def forward(ctx, pc_feats, img_feats, temperature, cmid_weight, gather):
        _lib.require_cuda(pc_feats, img_feats)
        pc = pc_feats.float().contiguous()
        n_r, D = pc.shape
        b = n_r // 2
        # packed rows: [t1; t2 | (t1+t2)/2; img]  -> one normalisation, one all-gather
        x = torch.empty((2 * n_r, D), dtype=F32, device=pc.device)
        _copy_rows(pc, x[:n_r])
        ops.add_scale(pc[:b], pc[b:], 0.5, out=x[n_r:n_r + b])
        _copy_rows(img_feats, x[n_r + b:])
        l, saved = _NTXentPack.fwd(x, 2, temperature, gather)
        total = torch.empty(3, dtype=F32, device=pc.device)
        ops.add_scale(l, None, 1.0, out=total[1:3])
        ops.add_scale(l[0:1], ops.add_scale(l[1:2], None, float(cmid_weight)), 1.0, out=total[0:1])
        ctx.saved = (saved, b, float(cmid_weight))
        return total

    @staticmethod
    def backward(ctx, dtotal):
        saved, b, w = ctx.saved
        up = dtotal.float().contiguous()[0:1]      # gradients flow through total[0] only (entries 1,2 are for logging)
        if _SERIAL_TERMS:
            d_imid, d_cmid = _NTXentPack.bwd(saved, 0, 1.0, up), _NTXentPack.bwd(saved, 1, w, up)
        else:
            d_all = _NTXentPack.bwd_all(saved, (1.0, w), up)          # one launch per stage for both terms
            d_imid, d_cmid = d_all[:2 * b], d_all[2 * b:]             # w.r.t. [t1; t2] and [(t1+t2)/2; img]
        half = ops.add_scale(d_cmid[:b], None, 0.5)
        dpc = torch.empty_like(d_imid)
        ops.add_scale(d_imid[:b], half, 1.0, out=dpc[:b])
        ops.add_scale(d_imid[b:], half, 1.0, out=dpc[b:])
        return dpc, d_cmid[b:], None, None, None


def pretrain_loss(pc_feats, img_feats, temperature=0.1, cmid_weight=1.0, gather_distributed=False):
    """-> fp32 tensor [3] = (total, loss_imid, loss_cmid); backpropagate through element 0."""
    return _PretrainLossFn.apply(pc_feats, img_feats, float(temperature), float(cmid_weight), bool(gather_distributed))
