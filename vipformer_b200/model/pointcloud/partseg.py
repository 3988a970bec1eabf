"""Drop-in mirror of the reference's vipformer/model/pointcloud/partseg.py (the `--mp` models).

Same class names, constructor arguments, forward signatures, error behaviour and state_dict keys as
the reference (MultiHeadAttention:15-86, CrossAttention:89-116, SelfAttention:119-141,
CrossAttentionLayer:144-167, SelfAttentionLayer:170-188, MLP:191-198, Residual:201-213,
Encoder:233-342, CrossFormer_pc_mp:473-550, CrossFormer_pc_mp_ft:553-605, CrossFormer_img_mp:608-680),
but every forward/backward runs the hand-written sm_100a kernels behind include/vpf.h via
`vipformer_b200.functional`.  The torch.nn leaf modules below (nn.Linear, nn.LayerNorm, ...) are used as
PARAMETER CONTAINERS ONLY -- their ATen forward is never called; there is no PyTorch/CPU fallback.

Granularity: the fused kernels sit at the layer level, so `Encoder`, `CrossAttentionLayer`,
`SelfAttentionLayer`, `Group2Emb`, `PointCloudInputAdapter`, the position / patch embeddings and the two
CrossFormer models are callable; calling an inner fragment on its own (MultiHeadAttention, CrossAttention,
SelfAttention, MLP, Residual) raises, because that fragment only exists inside a fused layer kernel sequence.
"""
from types import SimpleNamespace as NS
from typing import Optional

import torch
import torch.nn as nn

from ... import functional as Fn
from ... import functional_seg as FnS
from ... import _lib, ops, params
from .classifier import InputAdapter, PointCloudInputAdapter  # noqa: F401
from .utils import Group2Emb, PointNetFeaturePropagation, Sequential, divide_patches  # noqa: F401  (as the reference)

BF16, F32 = torch.bfloat16, torch.float32


from ... import runtime as _rt
from ...runtime import (StepState as _StepState, advance_dropout_seed, anchor as _anchor, as_f32_2d as _as_f32_2d,  # noqa: F401,E501
                        bump as _bump, fragment_error as _fragment_error, manual_seed, root_prepare as _root_prepare,
                        set_root as _set_root)


_TOK_STREAM = __import__("os").environ.get("VPF_TOK_STREAM", "1") != "0"


# ------------------------------------------------------------------------------------- parameter containers
class MultiHeadAttention(nn.Module):
    def __init__(self, num_heads: int, num_q_input_channels: int, num_kv_input_channels: int,
                 num_latent_channels: int, num_output_channels: Optional[int] = None, dropout: float = 0.0):
        super().__init__()
        if num_output_channels is None:
            num_output_channels = num_q_input_channels
        if num_latent_channels % num_heads != 0:
            raise ValueError("num_latent_channels must be divisible by num_heads")
        num_channels_per_head = num_latent_channels // num_heads
        self.dp_scale = num_channels_per_head ** -0.5
        self.num_heads = num_heads
        self.q_proj = nn.Linear(num_q_input_channels, num_latent_channels, bias=False)
        self.k_proj = nn.Linear(num_kv_input_channels, num_latent_channels, bias=False)
        self.v_proj = nn.Linear(num_kv_input_channels, num_latent_channels, bias=False)
        self.o_proj = nn.Linear(num_latent_channels, num_latent_channels)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x_q, x_kv, pad_mask=None, attn_mask=None):
        if attn_mask is not None:
            raise NotImplementedError("attention masks not supported yet")
        _fragment_error("MultiHeadAttention")


class CrossAttention(nn.Module):
    def __init__(self, num_heads: int, num_q_input_channels: int, num_kv_input_channels: int,
                 num_latent_channels: int, dropout: float = 0.0):
        super().__init__()
        self.q_norm = nn.LayerNorm(num_q_input_channels)
        self.kv_norm = nn.LayerNorm(num_kv_input_channels)
        self.attention = MultiHeadAttention(num_heads=num_heads, num_q_input_channels=num_q_input_channels,
                                            num_kv_input_channels=num_kv_input_channels,
                                            num_latent_channels=num_latent_channels,
                                            num_output_channels=num_latent_channels, dropout=dropout)

    def forward(self, x_q, x_kv, pad_mask=None, attn_mask=None):
        if attn_mask is not None:
            raise NotImplementedError("attention masks not supported yet")
        _fragment_error("CrossAttention")


class SelfAttention(nn.Module):
    def __init__(self, num_heads: int, num_latent_channels: int, dropout: float = 0.0):
        super().__init__()
        self.norm = nn.LayerNorm(num_latent_channels)
        self.attention = MultiHeadAttention(num_heads=num_heads, num_q_input_channels=num_latent_channels,
                                            num_kv_input_channels=num_latent_channels,
                                            num_latent_channels=num_latent_channels,
                                            num_output_channels=num_latent_channels, dropout=dropout)

    def forward(self, x, pad_mask=None, attn_mask=None):
        if attn_mask is not None:
            raise NotImplementedError("attention masks not supported yet")
        _fragment_error("SelfAttention")


class MLP(Sequential):
    def __init__(self, num_channels: int, widening_factor: int):
        super().__init__(nn.LayerNorm(num_channels), nn.Linear(num_channels, widening_factor * num_channels),
                         nn.GELU(), nn.Linear(widening_factor * num_channels, num_channels))

    def forward(self, *x):
        _fragment_error("MLP")


class Residual(nn.Module):
    def __init__(self, module: nn.Module, dropout: float, drop_path_rate: float):
        super().__init__()
        self.module = module
        self.dropout = nn.Dropout(p=dropout)
        # timm's DropPath has no parameters or buffers: an Identity placeholder keeps the state_dict identical; the rate is
        # applied by the enclosing fused layer (per-sample scales from the device step seed, functional.sa_layer_fwd)
        self.drop_path = nn.Identity()
        self.drop_path_rate = float(drop_path_rate)

    def forward(self, *args, **kwargs):
        _fragment_error("Residual")


# --------------------------------------------------------------------------------------------- fused layers
def _mlp_weights(mlp, W, G=None):
    ln, fc1, fc2 = mlp[0], mlp[1], mlp[3]
    if G is None:
        W.ln2_w, W.ln2_b = ln.weight, ln.bias
        W.w1, W.b1, W.w2, W.b2 = params.wb(fc1.weight), fc1.bias, params.wb(fc2.weight), fc2.bias
    else:
        G.ln2_w, G.ln2_b = ln.weight.grad, ln.bias.grad
        G.w1, G.b1, G.w2, G.b2 = fc1.weight.grad, fc1.bias.grad, fc2.weight.grad, fc2.bias.grad


class _LayerBase(Sequential):
    _op_counter = 0

    def _init_op_base(self):
        if self._D // self._H != 64:
            # the attention kernels are built for head_dim 64 (every published ViPFormer config: D256/H4, D384/H6)
            raise NotImplementedError(f"head dim {self._D // self._H} (= num_latent_channels / num_heads) is not built; "
                                      "the sm_100a attention kernels are specialised for 64")
        _LayerBase._op_counter += 8
        self._op_base = _LayerBase._op_counter

    def _p(self, training):
        return (self._p_attn, self._p_res1, self._p_res2) if training else (0.0, 0.0, 0.0)

    def _path(self, training):
        return float(getattr(self, "_p_path", 0.0)) if training else 0.0


class CrossAttentionLayer(_LayerBase):
    def __init__(self, num_heads: int, num_q_input_channels: int, num_kv_input_channels: int,
                 num_latent_channels: int, widening_factor: int = 1, drop_path_rate: float = 0.0,
                 atten_drop: float = 0.0, mlp_drop: float = 0.0, attention_residual: bool = True):
        if not attention_residual:
            raise NotImplementedError("attention_residual=False is not used by any ViPFormer model")
        if not (num_q_input_channels == num_kv_input_channels == num_latent_channels):
            raise NotImplementedError("fused cross-attention layer needs equal q/kv/latent channel counts")
        cross_attn = CrossAttention(num_heads=num_heads, num_q_input_channels=num_q_input_channels,
                                    num_kv_input_channels=num_kv_input_channels,
                                    num_latent_channels=num_latent_channels, dropout=atten_drop)
        super().__init__(Residual(cross_attn, atten_drop, drop_path_rate),
                         Residual(MLP(num_q_input_channels, widening_factor), mlp_drop, drop_path_rate))
        if drop_path_rate > 0.0:
            raise NotImplementedError("DropPath on the cross-attention layer is not built (Encoder never sets it, partseg.py:286-295)")
        # partseg.py:165-166: attention residual dropout = atten_drop, MLP residual dropout = mlp_drop
        self._p_attn, self._p_res1, self._p_res2 = atten_drop, atten_drop, mlp_drop
        self._H, self._D = num_heads, num_latent_channels
        self._init_op_base()

    def _weights(self):
        ca, mha = self[0].module, self[0].module.attention
        W = NS(qn_w=ca.q_norm.weight, qn_b=ca.q_norm.bias, kvn_w=ca.kv_norm.weight, kvn_b=ca.kv_norm.bias,
               wq=params.wb(mha.q_proj.weight),
               wkv=params.cat_view([params.wb(mha.k_proj.weight), params.wb(mha.v_proj.weight)]),
               wo=params.wb(mha.o_proj.weight), bo=mha.o_proj.bias)
        _mlp_weights(self[1].module, W)
        return W

    def _grads(self):
        ca, mha = self[0].module, self[0].module.attention
        G = NS(qn_w=ca.q_norm.weight.grad, qn_b=ca.q_norm.bias.grad, kvn_w=ca.kv_norm.weight.grad,
               kvn_b=ca.kv_norm.bias.grad, wq=mha.q_proj.weight.grad,
               wkv=params.cat_view([mha.k_proj.weight.grad, mha.v_proj.weight.grad]),
               wo=mha.o_proj.weight.grad, bo=mha.o_proj.bias.grad)
        _mlp_weights(self[1].module, None, G)
        return G

    def _cfg(self, B, L, Lk):
        pa, p1, p2 = self._p(self.training)
        return NS(B=B, L=L, Lk=Lk, D=self._D, H=self._H, scale=self[0].module.attention.dp_scale, p_attn=pa,
                  p_res1=p1, p_res2=p2)

    def forward(self, x_q, x_kv, pad_mask=None, attn_mask=None):
        """partseg.py:144-167 applied to (x_q, x_kv): x_q [B,L,D] fp32, x_kv [B,Lk,D] fp32 or bf16."""
        if attn_mask is not None:
            raise NotImplementedError("attention masks not supported yet")
        if pad_mask is not None:
            raise NotImplementedError("pad_mask is always None on the ViPFormer path and is not implemented")
        return _EncoderFn.apply(x_q, None, x_kv, _anchor(self), _SingleLayer(self, True))


class SelfAttentionLayer(_LayerBase):
    def __init__(self, num_heads: int, num_latent_channels: int, widening_factor: int = 1,
                 drop_path_rate: float = 0.0, atten_drop: float = 0.0, mlp_drop: float = 0.0):
        self_attn = SelfAttention(num_heads=num_heads, num_latent_channels=num_latent_channels, dropout=atten_drop)
        # partseg.py:186-187: BOTH residuals of a self-attention layer use mlp_drop
        super().__init__(Residual(self_attn, mlp_drop, drop_path_rate),
                         Residual(MLP(num_latent_channels, widening_factor), mlp_drop, drop_path_rate))
        self._p_attn, self._p_res1, self._p_res2 = atten_drop, mlp_drop, mlp_drop
        self._p_path = float(drop_path_rate)          # both Residuals of the layer, independent draws (partseg.py:186-187)
        self._H, self._D = num_heads, num_latent_channels
        self._init_op_base()

    def _weights(self):
        sa, mha = self[0].module, self[0].module.attention
        W = NS(ln1_w=sa.norm.weight, ln1_b=sa.norm.bias,
               wqkv=params.cat_view([params.wb(mha.q_proj.weight), params.wb(mha.k_proj.weight),
                                     params.wb(mha.v_proj.weight)]),
               wo=params.wb(mha.o_proj.weight), bo=mha.o_proj.bias)
        _mlp_weights(self[1].module, W)
        return W

    def _grads(self):
        sa, mha = self[0].module, self[0].module.attention
        G = NS(ln1_w=sa.norm.weight.grad, ln1_b=sa.norm.bias.grad,
               wqkv=params.cat_view([mha.q_proj.weight.grad, mha.k_proj.weight.grad, mha.v_proj.weight.grad]),
               wo=mha.o_proj.weight.grad, bo=mha.o_proj.bias.grad)
        _mlp_weights(self[1].module, None, G)
        return G

    def _cfg(self, B, L, Lk=None):
        pa, p1, p2 = self._p(self.training)
        return NS(B=B, L=L, Lk=L, D=self._D, H=self._H, scale=self[0].module.attention.dp_scale, p_attn=pa,
                  p_res1=p1, p_res2=p2, p_path=self._path(self.training))

    def forward(self, x, pad_mask=None, attn_mask=None):
        """partseg.py:170-188 applied to x [B,L,D] fp32."""
        if attn_mask is not None:
            raise NotImplementedError("attention masks not supported yet")
        if pad_mask is not None:
            raise NotImplementedError("pad_mask is always None on the ViPFormer path and is not implemented")
        return _EncoderFn.apply(x, None, None, _anchor(self), _SingleLayer(self, False))


class _SingleLayer:
    """Adapter that lets one layer run through the Encoder autograd function."""

    def __init__(self, layer, is_cross):
        self.cross_attn_1 = layer if is_cross else None
        self.sa_layers = [] if is_cross else [layer]
        self.root = layer


class _EncoderFn(torch.autograd.Function):
    """cross_attn_1 (optional) followed by the self-attention stack, positional term re-added before every layer
    (Encoder.forward, partseg.py:314-342).  One autograd node; explicit kernel-sequence backward."""

    @staticmethod
    def forward(ctx, x_q, pos, kv, anchor, enc, taps=None):
        """taps: None -> the last layer's output; a tuple of 1-based self-attention layer numbers -> the outputs of
        exactly those layers (Encoder.forward with modal_prior=False, partseg.py:328-340), as a tuple."""
        _lib.require_cuda(x_q)
        root = getattr(enc, "root", enc)
        arena = _root_prepare(root, x_q.device)
        B, L, D = x_q.shape
        x = _as_f32_2d(x_q, D)
        pos2 = None if pos is None else _as_f32_2d(pos, D)
        if pos2 is not None and pos2.shape[0] not in (B * L, L):
            raise ValueError(f"pos_embs must be [B,{L},{D}] or [1,{L},{D}]")
        save = any(ctx.needs_input_grad)
        seed = _StepState.seed_ptr(x_q.device)
        ctxs = []
        ca = enc.cross_attn_1
        training = (ca is not None and ca.training) or any(l.training for l in enc.sa_layers)
        off = _rt.next_op_offset(arena.managed) if training else 0
        if ca is not None:
            Lk = kv.shape[1]
            kv2 = kv.reshape(-1, D)
            if kv2.dtype not in (BF16, F32):
                kv2 = kv2.float()
            kv2 = kv2.contiguous()
            cfg = ca._cfg(B, L, Lk)
            x, c = Fn.ca_layer_fwd(x, pos2, kv2, ca._weights(), cfg, seed, ca._op_base + off, save)
            ctxs.append((ca, cfg, c, ca._op_base + off))
        outs, tap_at = [], []
        n_run = len(enc.sa_layers) if taps is None else (max(taps) if taps else 0)
        for i, layer in enumerate(enc.sa_layers):
            if i >= n_run:       # layers behind the last tap do not reach any output (the reference runs them for nothing)
                break
            cfg = layer._cfg(B, L)
            x, c = Fn.sa_layer_fwd(x, pos2, layer._weights(), cfg, seed, layer._op_base + off, save)
            ctxs.append((layer, cfg, c, layer._op_base + off))
            if taps is not None and (i + 1) in taps:
                outs.append(x.view(B, L, D))
                tap_at.append(len(ctxs) - 1)
        ctx.tap_at = tap_at if taps is not None else None
        if save:
            ctx.ctxs, ctx.arena, ctx.has_pos = ctxs, arena, pos2 is not None
            ctx.pos_shape = None if pos is None else pos.shape
            ctx.kv_shape = None if kv is None else kv.shape
            ctx.kv_dtype = None if kv is None else kv.dtype
            ctx.seed, ctx.shape = seed, (B, L, D)
        if taps is not None:
            return tuple(outs)
        return x.view(B, L, D)

    @staticmethod
    def backward(ctx, *douts):
        B, L, D = ctx.shape
        ctx.arena.ensure_grads()
        dev = next(d for d in douts if d is not None).device
        tap_grads = {}
        if ctx.tap_at is None:
            dx = _as_f32_2d(douts[0], D)
        else:        # gradient of tap j enters the chain behind the layer that produced it
            dx = None
            for pos_i, d in zip(ctx.tap_at, douts):
                if d is not None:
                    tap_grads[pos_i] = _as_f32_2d(d, D)
        dpos = None
        if ctx.has_pos:
            dpos = ops.zeros_(torch.empty(ctx.pos_shape, dtype=F32, device=dev))
        dkv = None
        g = None     # masked bf16 copy of dx, emitted by the layer above for this layer's MLP residual (functional._ln_bwd)
        for li in range(len(ctx.ctxs) - 1, -1, -1):
            layer, cfg, c, op_base = ctx.ctxs[li]
            if li in tap_grads:
                dx = tap_grads[li] if dx is None else ops.add_scale(dx, tap_grads[li], 1.0)
                g = None         # a tap gradient joined the stream here: the emitted copy no longer matches dx
            if dx is None:       # nothing flows into this layer (no tap at or behind it)
                continue
            G = layer._grads()
            if isinstance(layer, CrossAttentionLayer):
                dx, dkv = Fn.ca_layer_bwd(dx, c, layer._weights(), G, cfg, ctx.seed, op_base, dpos, g2=g)
                g = None
            else:
                emit = None
                if li > 0 and (li - 1) not in tap_grads and getattr(ctx.ctxs[li - 1][1], "p_path", 0.0) == 0.0:
                    # the layer below consumes dx through its MLP-residual dropout (unless its DropPath rescales dx first)
                    below, cfg_b, _, op_b = ctx.ctxs[li - 1]
                    emit = (cfg_b.p_res2, ctx.seed, op_b + 2, below._grads().b2)
                dx, g = Fn.sa_layer_bwd(dx, c, layer._weights(), G, cfg, ctx.seed, op_base, dpos, g2=g, emit=emit)
        ctx.ctxs = None
        if dkv is not None:
            dkv = dkv.view(ctx.kv_shape)
            if not ctx.needs_input_grad[2]:
                dkv = None
        return dx.view(B, L, D), dpos, dkv, None, None, None


class Encoder(nn.Module):
    def __init__(self, num_latent_channels: int, num_cross_attention_layers: int = 1,
                 num_cross_attention_heads: int = 4, cross_attention_widening_factor: int = 1,
                 first_cross_attention_layer_shared: bool = False, num_self_attention_layers: int = 6,
                 num_self_attention_heads: int = 4, self_attention_widening_factor: int = 1, dpr_list: list = [],
                 atten_drop: float = 0.0, mlp_drop: float = 0.0, activation_checkpointing: bool = False,
                 modal_prior: bool = False):
        super().__init__()
        if num_cross_attention_layers <= 0:
            raise ValueError("num_cross_attention_layers must be > 0")
        if num_cross_attention_layers != 1:
            raise NotImplementedError("only num_cross_attention_layers == 1 (every published ViPFormer config) is built")
        if activation_checkpointing:
            raise NotImplementedError("activation_checkpointing is not needed at 180 GB HBM and is not implemented")
        self.num_cross_attention_layers = num_cross_attention_layers

        def cross_attn():
            return CrossAttentionLayer(num_heads=num_cross_attention_heads, num_q_input_channels=num_latent_channels,
                                       num_kv_input_channels=num_latent_channels,
                                       num_latent_channels=num_latent_channels,
                                       widening_factor=cross_attention_widening_factor, atten_drop=atten_drop,
                                       mlp_drop=mlp_drop)

        self.cross_attn_n = cross_attn()
        # partseg.py:295-300: with one cross-attention layer both names alias the same module (and state_dict keys)
        self.cross_attn_1 = self.cross_attn_n
        self.sa_layers = nn.ModuleList()
        for i in range(num_self_attention_layers):
            self.sa_layers.append(SelfAttentionLayer(num_heads=num_self_attention_heads,
                                                     num_latent_channels=num_latent_channels,
                                                     widening_factor=self_attention_widening_factor,
                                                     drop_path_rate=dpr_list[i], atten_drop=atten_drop,
                                                     mlp_drop=mlp_drop))
        self.modal_prior = modal_prior

    def forward(self, group_embs, pos_embs, pts_embs, layer_idx=[], pad_mask=None):
        """partseg.py:314-342.  group_embs [B,G,D], pos_embs [B,G,D] or [1,G,D], pts_embs [B,N,D] -> [B,G,D]."""
        if pad_mask is not None:
            raise NotImplementedError("pad_mask is always None on the ViPFormer path and is not implemented")
        if not self.modal_prior:     # partseg.py:336-340: the outputs of the self-attention layers named in layer_idx
            taps = tuple(int(i) for i in layer_idx)
            if not taps or min(taps) < 1 or max(taps) > len(self.sa_layers) or list(taps) != sorted(set(taps)):
                raise ValueError("layer_idx must be a strictly increasing list of self-attention layer numbers (1-based)")
            return list(_EncoderFn.apply(group_embs, pos_embs, pts_embs, _anchor(self), self, taps))
        return _EncoderFn.apply(group_embs, pos_embs, pts_embs, _anchor(self), self)


# ------------------------------------------------------------------------------- embeddings, pooling, heads
class _PositionEmb(nn.Sequential):
    """nn.Sequential(Linear(3,128), GELU, Linear(128,D)) of partseg.py:498-501 with a fused forward."""

    def forward(self, center):
        return _PosEmbFn.apply(center, _anchor(self), self)


class _PosEmbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, center, anchor, mod):
        _lib.require_cuda(center)
        arena = _root_prepare(mod, center.device)
        center = center.float().contiguous()
        W = NS(w1=mod[0].weight, b1=mod[0].bias, w2=params.wb(mod[2].weight), b2=mod[2].bias)
        save = any(ctx.needs_input_grad)
        pos, c = Fn.posemb_fwd(center, W, save)
        if save:
            ctx.c, ctx.mod, ctx.arena, ctx.W = c, mod, arena, W
        return pos.view(center.shape[0], center.shape[1], -1)

    @staticmethod
    def backward(ctx, dpos):
        ctx.arena.ensure_grads()
        m = ctx.mod
        G = NS(w1=m[0].weight.grad, b1=m[0].bias.grad, w2=m[2].weight.grad, b2=m[2].bias.grad)
        Fn.posemb_bwd(_as_f32_2d(dpos, dpos.shape[-1]), ctx.c, ctx.W, G)
        ctx.c = None
        return None, None, None


class _Patch2Emb(nn.Sequential):
    """nn.Sequential(Rearrange('b (h p1) (w p2) c -> b (h w) (p1 p2 c)'), Linear) of partseg.py:631-634."""

    def __init__(self, patch_size, in_dim, out_dim):
        super().__init__(nn.Identity(), nn.Linear(in_dim, out_dim))   # index 1 keeps the key `patch2emb.1.weight`
        self.patch_size = patch_size

    def forward(self, imgs):
        return _Patch2EmbFn.apply(imgs, _anchor(self), self)


class _Patch2EmbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imgs, anchor, mod):
        _lib.require_cuda(imgs)
        arena = _root_prepare(mod, imgs.device)
        # a permuted NCHW view (pretrain.py:179 `torch.permute(imgs, (0, 2, 3, 1))`) is consumed in place
        nchw = (not imgs.is_contiguous()) and imgs.permute(0, 3, 1, 2).is_contiguous()
        imgs = imgs.float() if nchw else imgs.float().contiguous()
        W = NS(w=params.wb(mod[1].weight), b=mod[1].bias)
        save = any(ctx.needs_input_grad)
        e, c = Fn.patch2emb_fwd(imgs.permute(0, 3, 1, 2) if nchw else imgs, W, mod.patch_size, save, nchw)
        if save:
            ctx.c, ctx.mod, ctx.arena = c, mod, arena
        return e.view(imgs.shape[0], -1, e.shape[-1])

    @staticmethod
    def backward(ctx, de):
        ctx.arena.ensure_grads()
        G = NS(w=ctx.mod[1].weight.grad, b=ctx.mod[1].bias.grad)
        Fn.patch2emb_bwd(_as_f32_2d(de, de.shape[-1]), ctx.c, G)
        ctx.c = None
        return None, None, None


class _LatentHead(nn.Sequential):
    """latent_head of partseg.py:519-525 (BN1d, ReLU, Linear no-bias, BN1d, ReLU, Linear no-bias), fused with the
    max/mean token pooling of partseg.py:547."""

    def forward(self, x_latent):
        """x_latent [B,L,D] -> (feats [B,D], backbone [B,2D])."""
        return _PoolHeadFn.apply(x_latent, _anchor(self), self)


class _PoolHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, mod):
        _lib.require_cuda(x)
        arena = _root_prepare(mod, x.device)
        B, L, D = x.shape
        W = NS(bn1_w=mod[0].weight, bn1_b=mod[0].bias, wa=params.wb(mod[2].weight), bn2_w=mod[3].weight,
               bn2_b=mod[3].bias, wb=params.wb(mod[5].weight))
        bn = NS(rm1=mod[0].running_mean, rv1=mod[0].running_var, rm2=mod[3].running_mean, rv2=mod[3].running_var)
        save = any(ctx.needs_input_grad)
        feats, pooled, c = Fn.pool_head_fwd(_as_f32_2d(x, D), W, bn, B, L, D, mod.training, save)
        _rt.tap("head", c)
        if mod.training:
            _bump(mod[0]); _bump(mod[3])
        if save:
            ctx.c, ctx.mod, ctx.arena, ctx.W, ctx.shape = c, mod, arena, W, (B, L, D)
        return feats, pooled

    @staticmethod
    def backward(ctx, dfeats, dbackbone):
        ctx.arena.ensure_grads()
        m = ctx.mod
        B, L, D = ctx.shape
        G = NS(bn1_w=m[0].weight.grad, bn1_b=m[0].bias.grad, wa=m[2].weight.grad, bn2_w=m[3].weight.grad,
               bn2_b=m[3].bias.grad, wb=m[5].weight.grad)
        dfeats = None if dfeats is None else dfeats.float().contiguous()
        dbackbone = None if dbackbone is None else dbackbone.float().contiguous()
        dx = Fn.pool_head_bwd(dfeats, dbackbone, ctx.c, ctx.W, G, B, L, D)
        ctx.c = None
        return dx.view(B, L, D), None, None


def _latent_head(D):
    return _LatentHead(nn.BatchNorm1d(2 * D), nn.ReLU(), nn.Linear(2 * D, D, bias=False), nn.BatchNorm1d(D),
                       nn.ReLU(), nn.Linear(D, D, bias=False))


# -------------------------------------------------------------------------------------------------- models
class CrossFormer_pc_mp(nn.Module):
    """partseg.py:473-550.  forward(pts [B,N,3]) -> (feats [B,D], backbone [B,2D])."""

    def __init__(self, input_adapter=None, num_latents=128, num_latent_channels=384, group_size=32,
                 num_cross_attention_layers=1, num_cross_attention_heads=6, num_self_attention_layers=6,
                 num_self_attention_heads=6, mlp_widen_factor=4, max_dpr=.0, atten_drop=0.1, mlp_drop=.5,
                 modal_prior=True):
        super().__init__()
        self.num_groups = num_latents
        self.group_size = group_size
        self.group2emb = Group2Emb(num_latent_channels)
        self.position_emb = _PositionEmb(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, num_latent_channels))
        self.input_adapter = input_adapter
        dpr_list = [x.item() for x in torch.linspace(0, max_dpr, num_self_attention_layers)]
        self.encoder = Encoder(num_latent_channels=num_latent_channels,
                               num_cross_attention_layers=num_cross_attention_layers,
                               num_cross_attention_heads=num_cross_attention_heads,
                               cross_attention_widening_factor=mlp_widen_factor,
                               num_self_attention_layers=num_self_attention_layers,
                               num_self_attention_heads=num_self_attention_heads,
                               self_attention_widening_factor=mlp_widen_factor, dpr_list=dpr_list,
                               atten_drop=atten_drop, mlp_drop=mlp_drop, modal_prior=modal_prior)
        self.latent_head = _latent_head(num_latent_channels)
        # deterministic FPS hook (the reference draws torch.randint from the global RNG, utils.py:71)
        self.fps_start_idx = None
        self.fps_generator = None

    def _tokens(self, pts):
        if self.__dict__.get("_vpf_root") is None:
            _set_root(self)
        pts = pts.float().contiguous()
        if _TOK_STREAM and pts.is_cuda:
            # FPS / kNN are instruction-issue bound and touch ~1 % of the HBM bandwidth, the input adapter (point-wise MLP over
            # every point) is the opposite: run them side by side, the tokenizer on an auxiliary stream
            cur, aux = torch.cuda.current_stream(), _rt.aux_stream(pts.device)
            start = self.fps_start_idx
            aux.wait_stream(cur)
            with torch.cuda.stream(aux):
                neighborhood, center = divide_patches(pts, self.num_groups, self.group_size, start_idx=start,
                                                      generator=self.fps_generator)
            pts_embs = self.input_adapter(pts)
            cur.wait_stream(aux)
        else:
            pts_embs = self.input_adapter(pts)
            neighborhood, center = divide_patches(pts, self.num_groups, self.group_size, start_idx=self.fps_start_idx,
                                                  generator=self.fps_generator)
        group_embs = self.group2emb(neighborhood)
        pos_embs = self.position_emb(center)
        return self.encoder(group_embs, pos_embs, pts_embs)

    def forward(self, pts):
        x_latent = self._tokens(pts)
        x_latent_feats, backbone_feats = self.latent_head(x_latent)
        return x_latent_feats, backbone_feats


class _FinetuneHead(nn.Sequential):
    """finetune_head of partseg.py:573-582: 3 x {BatchNorm1d, ReLU, Linear (with bias)}, fused with the max/mean token
    pooling of partseg.py:601.  forward(x_latent [B,L,D]) -> logits [B, num_obj_classes]."""

    def _stages(self, grads=False):
        out = []
        for bn_i, fc_i in ((0, 2), (3, 5), (6, 8)):
            bn, fc = self[bn_i], self[fc_i]
            if grads:
                out.append(NS(bn_w=bn.weight.grad, bn_b=bn.bias.grad, w=fc.weight.grad, b=fc.bias.grad))
            else:
                out.append(NS(bn_w=bn.weight, bn_b=bn.bias, rm=bn.running_mean, rv=bn.running_var,
                              w=params.wb(fc.weight), b=fc.bias))
        return out

    def forward(self, x_latent):
        return _PoolClsHeadFn.apply(x_latent, _anchor(self), self)


class _PoolClsHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, mod):
        _lib.require_cuda(x)
        arena = _root_prepare(mod, x.device)
        B, L, D = x.shape
        stages = mod._stages()
        save = any(ctx.needs_input_grad)
        logits, c = Fn.pool_cls_head_fwd(_as_f32_2d(x, D), stages, B, L, D, mod.training, save)
        _rt.tap("cls_head", c)
        if mod.training:
            _bump(mod[0]); _bump(mod[3]); _bump(mod[6])
        if save:
            ctx.c, ctx.mod, ctx.arena, ctx.stages, ctx.shape = c, mod, arena, stages, (B, L, D)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.arena.ensure_grads()
        B, L, D = ctx.shape
        dx = Fn.pool_cls_head_bwd(dlogits.float(), ctx.c, ctx.stages, ctx.mod._stages(grads=True), B, L, D)
        ctx.c = None
        return dx.view(B, L, D), None, None


class CrossFormer_pc_mp_ft(CrossFormer_pc_mp):
    """partseg.py:553-605: the pre-trained point-cloud branch + a 3-stage classification head; forward(pts [B,N,3]) ->
    logits [B, num_obj_classes].  `latent_head` stays in the module (and in the state_dict) unused, exactly as in the
    reference, so pre-training checkpoints load with strict=False and only `finetune_head.*` is missing (ft_cls.py:92-98)."""

    def __init__(self, input_adapter=None, num_latents=128, num_latent_channels=384, group_size=32,
                 num_cross_attention_layers=1, num_cross_attention_heads=6, num_self_attention_layers=6,
                 num_self_attention_heads=6, mlp_widen_factor=4, max_dpr=0, atten_drop=0.1, mlp_drop=0.5,
                 modal_prior=True, num_obj_classes=40):
        super().__init__(input_adapter, num_latents, num_latent_channels, group_size, num_cross_attention_layers,
                         num_cross_attention_heads, num_self_attention_layers, num_self_attention_heads,
                         mlp_widen_factor, max_dpr, atten_drop, mlp_drop, modal_prior)
        D = num_latent_channels
        self.finetune_head = _FinetuneHead(nn.BatchNorm1d(2 * D), nn.ReLU(), nn.Linear(2 * D, D), nn.BatchNorm1d(D), nn.ReLU(),
                                           nn.Linear(D, D // 2), nn.BatchNorm1d(D // 2), nn.ReLU(),
                                           nn.Linear(D // 2, num_obj_classes))

    def forward(self, pts):
        return self.finetune_head(self._tokens(pts))


class _PartSegHeadFn(torch.autograd.Function):
    """Everything of CrossFormer_partseg.forward behind the encoder taps (partseg.py:420-468) as one autograd node."""

    @staticmethod
    def forward(ctx, pts, center, cls_label, anchor, mod, *taps):
        _lib.require_cuda(pts)
        arena = _root_prepare(mod, pts.device)
        B, N, _ = pts.shape
        G, D = taps[0].shape[1], taps[0].shape[2]
        cfg = NS(B=B, N=N, G=G, D=D, P=mod.conv3.out_channels, p_dp1=mod.dp1.p)
        W, bn = mod._head_weights()
        save = any(ctx.needs_input_grad)
        seed = _StepState.seed_ptr(pts.device)
        off = _rt.next_op_offset(arena.managed) if mod.training else 0
        logits, c = FnS.partseg_head_fwd([_as_f32_2d(t, D) for t in taps], pts.float().contiguous(), center.float().contiguous(),
                                         cls_label.reshape(B, -1), W, bn, cfg, mod.training, seed, mod._op_base + off, save)
        _rt.tap("seg_head", c)
        if mod.training:
            for m in (mod.label_conv[1], mod.propagation.mlp_bns[0], mod.propagation.mlp_bns[1], mod.bn1, mod.bn2):
                _bump(m)
        if save:
            ctx.c, ctx.mod, ctx.arena, ctx.W, ctx.cfg, ctx.seed, ctx.op = c, mod, arena, W, cfg, seed, mod._op_base + off
        return logits.view(B, N, cfg.P)

    @staticmethod
    def backward(ctx, dlogits):
        ctx.arena.ensure_grads()
        cfg = ctx.cfg
        d2 = dlogits.float().reshape(cfg.B * cfg.N, cfg.P)
        if d2.stride(1) != 1:
            d2 = d2.contiguous()
        dtaps = FnS.partseg_head_bwd(d2, ctx.c, ctx.W, ctx.mod._head_grads(), cfg, ctx.seed, ctx.op)
        ctx.c = None
        return (None, None, None, None, None) + tuple(d.view(cfg.B, cfg.G, cfg.D) for d in dtaps)


class CrossFormer_partseg(nn.Module):
    """partseg.py:345-470.  forward(pts [B,N,3], cls_label [B,16]) -> part logits [B, N, num_part_classes].
    DropPath (max_dpr > 0, the reference's default 0.1) is applied by the fused self-attention layers."""

    def __init__(self, input_adapter=None, num_latents=128, num_latent_channels=384, group_size=32,
                 num_cross_attention_layers=1, num_cross_attention_heads=6, num_self_attention_layers=12,
                 num_self_attention_heads=6, mlp_widen_factor=4, max_dpr=0.1, atten_drop=.0, mlp_drop=.0, layer_idx=[],
                 num_part_classes=50):
        super().__init__()
        D = num_latent_channels
        self.num_groups = num_latents
        self.group_size = group_size
        self.group2emb = Group2Emb(D)
        self.position_emb = _PositionEmb(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, D))
        self.input_adapter = input_adapter
        dpr_list = [x.item() for x in torch.linspace(0, max_dpr, num_self_attention_layers)]
        self.encoder = Encoder(num_latent_channels=D, num_cross_attention_layers=num_cross_attention_layers,
                               num_cross_attention_heads=num_cross_attention_heads,
                               cross_attention_widening_factor=mlp_widen_factor,
                               num_self_attention_layers=num_self_attention_layers,
                               num_self_attention_heads=num_self_attention_heads,
                               self_attention_widening_factor=mlp_widen_factor, dpr_list=dpr_list, atten_drop=atten_drop,
                               mlp_drop=mlp_drop)
        self.layer_idx = layer_idx
        self.norm = nn.LayerNorm(D)
        self.label_conv = nn.Sequential(nn.Conv1d(16, 64, kernel_size=1, bias=False), nn.BatchNorm1d(64), nn.LeakyReLU(0.2))
        self.num_layer_idx = len(layer_idx)
        self.propagation = PointNetFeaturePropagation(in_channel=self.num_layer_idx * D + 3, mlp=[mlp_widen_factor * D, 1024])
        self.conv1 = nn.Conv1d(2 * self.num_layer_idx * D + 64 + 1024, 512, 1)
        self.bn1 = nn.BatchNorm1d(512)
        self.dp1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(512, 256, 1)
        self.bn2 = nn.BatchNorm1d(256)
        self.conv3 = nn.Conv1d(256, num_part_classes, 1)
        self.relu = nn.ReLU()
        self.fps_start_idx = None
        self.fps_generator = None
        _LayerBase._op_counter += 8
        self._op_base = _LayerBase._op_counter       # dropout stream of dp1

    def _head_weights(self):
        wb, pr = params.wb, self.propagation
        k, D = self.num_layer_idx, self.norm.normalized_shape[0]
        W = NS(ln_w=self.norm.weight, ln_b=self.norm.bias, lc_w=wb(self.label_conv[0].weight).view(64, 16),
               lc_bn_w=self.label_conv[1].weight, lc_bn_b=self.label_conv[1].bias,
               p1_w_f32=pr.mlp_convs[0].weight.view(-1, k * D + 3), p1_b=pr.mlp_convs[0].bias,
               p1_bn_w=pr.mlp_bns[0].weight, p1_bn_b=pr.mlp_bns[0].bias,
               p2_w=wb(pr.mlp_convs[1].weight).view(1024, -1), p2_b=pr.mlp_convs[1].bias,
               p2_bn_w=pr.mlp_bns[1].weight, p2_bn_b=pr.mlp_bns[1].bias,
               c1_w=wb(self.conv1.weight).view(512, -1), c1_b=self.conv1.bias, bn1_w=self.bn1.weight, bn1_b=self.bn1.bias,
               c2_w=wb(self.conv2.weight).view(256, 512), c2_b=self.conv2.bias, bn2_w=self.bn2.weight, bn2_b=self.bn2.bias,
               c3_w=wb(self.conv3.weight).view(-1, 256), c3_b=self.conv3.bias)
        bn = NS(lc_rm=self.label_conv[1].running_mean, lc_rv=self.label_conv[1].running_var,
                p1_rm=pr.mlp_bns[0].running_mean, p1_rv=pr.mlp_bns[0].running_var, p2_rm=pr.mlp_bns[1].running_mean,
                p2_rv=pr.mlp_bns[1].running_var, rm1=self.bn1.running_mean, rv1=self.bn1.running_var,
                rm2=self.bn2.running_mean, rv2=self.bn2.running_var)
        return W, bn

    def _head_grads(self):
        pr = self.propagation
        k, D = self.num_layer_idx, self.norm.normalized_shape[0]
        return NS(ln_w=self.norm.weight.grad, ln_b=self.norm.bias.grad, lc_w=self.label_conv[0].weight.grad.view(64, 16),
                  lc_bn_w=self.label_conv[1].weight.grad, lc_bn_b=self.label_conv[1].bias.grad,
                  p1_w=pr.mlp_convs[0].weight.grad.view(-1, k * D + 3), p1_b=pr.mlp_convs[0].bias.grad,
                  p1_bn_w=pr.mlp_bns[0].weight.grad, p1_bn_b=pr.mlp_bns[0].bias.grad,
                  p2_w=pr.mlp_convs[1].weight.grad.view(1024, -1), p2_b=pr.mlp_convs[1].bias.grad,
                  p2_bn_w=pr.mlp_bns[1].weight.grad, p2_bn_b=pr.mlp_bns[1].bias.grad,
                  c1_w=self.conv1.weight.grad.view(512, -1), c1_b=self.conv1.bias.grad, bn1_w=self.bn1.weight.grad,
                  bn1_b=self.bn1.bias.grad, c2_w=self.conv2.weight.grad.view(256, 512), c2_b=self.conv2.bias.grad,
                  bn2_w=self.bn2.weight.grad, bn2_b=self.bn2.bias.grad, c3_w=self.conv3.weight.grad.view(-1, 256),
                  c3_b=self.conv3.bias.grad)

    def forward(self, pts, cls_label):
        if self.__dict__.get("_vpf_root") is None:
            _set_root(self)
        if self.num_layer_idx < 1:
            raise ValueError("CrossFormer_partseg needs a non-empty layer_idx")
        pts = pts.float().contiguous()
        pts_embs = self.input_adapter(pts)
        neighborhood, center = divide_patches(pts, self.num_groups, self.group_size, start_idx=self.fps_start_idx,
                                              generator=self.fps_generator)
        group_embs = self.group2emb(neighborhood)
        pos_embs = self.position_emb(center)
        feature_list = self.encoder(group_embs, pos_embs, pts_embs, self.layer_idx)
        return _PartSegHeadFn.apply(pts, center, cls_label, _anchor(self.propagation), self, *feature_list)


def load_pretrained(model, state_dict):
    """ft_cls.py:92-98 without the DDP wrapper: load a pre-training checkpoint of CrossFormer_pc_mp into a fine-tune model
    (`strict=False`: `finetune_head.*` stays at its initialisation).  Keys saved from a DDP-wrapped model ("module.") are
    accepted.  Returns the (missing, unexpected) key lists of load_state_dict."""
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    res = model.load_state_dict(sd, strict=False)
    return list(res.missing_keys), list(res.unexpected_keys)


class CrossFormer_img_mp(nn.Module):
    """partseg.py:608-680.  forward(imgs [B,H,W,3] NHWC) -> (feats [B,D], backbone [B,2D])."""

    def __init__(self, img_height=144, img_width=144, patch_size=12, num_latent_channels=384,
                 num_cross_attention_layers=1, num_cross_attention_heads=6, num_self_attention_layers=6,
                 num_self_attention_heads=6, mlp_widen_factor=4, max_dpr=.0, atten_drop=0.1, mlp_drop=.5,
                 modal_prior=True):
        super().__init__()
        num_patches = (img_height // patch_size) * (img_width // patch_size)
        self.patch2emb = _Patch2Emb(patch_size, patch_size * patch_size * 3, num_latent_channels)
        self.position_emb = nn.Parameter(torch.randn(1, num_patches, num_latent_channels))
        dpr_list = [x.item() for x in torch.linspace(0, max_dpr, num_self_attention_layers)]
        self.encoder = Encoder(num_latent_channels=num_latent_channels,
                               num_cross_attention_layers=num_cross_attention_layers,
                               num_cross_attention_heads=num_cross_attention_heads,
                               cross_attention_widening_factor=mlp_widen_factor,
                               num_self_attention_layers=num_self_attention_layers,
                               num_self_attention_heads=num_self_attention_heads,
                               self_attention_widening_factor=mlp_widen_factor, dpr_list=dpr_list,
                               atten_drop=atten_drop, mlp_drop=mlp_drop, modal_prior=modal_prior)
        self.latent_head = _latent_head(num_latent_channels)

    def forward(self, imgs):
        if self.__dict__.get("_vpf_root") is None:
            _set_root(self)
        patch_embs = self.patch2emb(imgs)
        pos_embs = self.position_emb
        x_latent = self.encoder(patch_embs, pos_embs, patch_embs)
        x_latent_feats, backbone_feats = self.latent_head(x_latent)
        return x_latent_feats, backbone_feats
