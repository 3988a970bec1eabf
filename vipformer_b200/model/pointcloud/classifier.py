"""Mirror of vipformer/model/pointcloud/classifier.py:25-50 (PointCloudInputAdapter) and of the InputAdapter base
(vipformer/model/core/modules.py, re-declared in partseg.py:216-230 of the reference)."""
from types import SimpleNamespace as NS
from typing import Tuple

import torch
import torch.nn as nn

from ... import functional as Fn
from ... import _lib, params
from ... import runtime as rt


class InputAdapter(nn.Module):
    def __init__(self, num_input_channels: int):
        super().__init__()
        self._num_input_channels = num_input_channels

    @property
    def num_input_channels(self):
        return self._num_input_channels

    def forward(self, x):
        raise NotImplementedError()


class PointCloudInputAdapter(InputAdapter):
    """Linear(C,64) -> LayerNorm(64) -> ReLU -> Linear(64,D) per point.  forward(x [B,N,C]) -> [B,N,D] (bf16:
    the result only feeds the cross-attention K/V path, which consumes bf16)."""

    def __init__(self, pointcloud_shape: Tuple[int, ...], num_input_channels: int):
        super().__init__(num_input_channels=num_input_channels)
        _, self.point_channels = pointcloud_shape
        if self.point_channels != 3:
            raise NotImplementedError("PointCloudInputAdapter kernels are built for xyz input (point_channels == 3)")
        self.point_mlp = nn.Sequential(nn.Linear(self.point_channels, 64), nn.LayerNorm(64), nn.ReLU(),
                                       nn.Linear(64, num_input_channels))

    def forward(self, x):
        return _AdapterFn.apply(x, rt.anchor(self), self)


class _AdapterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, anchor, mod):
        _lib.require_cuda(pts)
        arena = rt.root_prepare(mod, pts.device)
        pts = pts.float().contiguous()
        m = mod.point_mlp
        W = NS(w1=m[0].weight, b1=m[0].bias, ln_w=m[1].weight, ln_b=m[1].bias, w2=params.wb(m[3].weight), b2=m[3].bias)
        save = any(ctx.needs_input_grad)
        e, c = Fn.adapter_fwd(pts, W, save)
        rt.tap("adapter", c)
        if save:
            ctx.c, ctx.mod, ctx.arena, ctx.W = c, mod, arena, W
        return e.view(pts.shape[0], pts.shape[1], -1)

    @staticmethod
    def backward(ctx, de):
        ctx.arena.ensure_grads()
        m = ctx.mod.point_mlp
        G = NS(w1=m[0].weight.grad, b1=m[0].bias.grad, ln_w=m[1].weight.grad, ln_b=m[1].bias.grad,
               w2=m[3].weight.grad, b2=m[3].bias.grad)
        Fn.adapter_bwd(de.reshape(-1, de.shape[-1]).contiguous(), ctx.c, ctx.W, G)
        ctx.c = None
        return None, None, None
