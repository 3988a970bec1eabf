"""Device-side counterpart of the reference's point-cloud input pipeline (datasets/data.py:16-36,92-118,
datasets/data_utils.py:56-221): the augmentation chain trans_1 / trans_2 as one CUDA kernel over a batch of clouds.

    aug = DeviceAugment()                       # Normalize, Scale(0.5, 2), Rotate(y), Translate(0.5), Jitter(0.01, 0.05), Dropout(0.875)
    pc_t1, pc_t2 = aug(raw), aug(raw)           # raw [B, N, 3] fp32 on the GPU: two independently augmented views (data.py:109-112)

At > 10 k shapes/s per GPU the reference's 18 CPU workers (numpy + PLY parsing) cannot feed the step; the file parsing stays
on the host, the arithmetic moves here.  Random draws come from a torch device generator (plumbing) and are handed to the
kernel as inputs, so `draws=` reproduces any given sample exactly (tests, oracle/augment.py).
"""
import ctypes
import math

import torch

from .. import _lib

F32 = torch.float32


def draw(B, N, device, generator=None, lo=0.5, hi=2.0, translate_range=0.5, jitter_std=0.01, max_dropout_ratio=0.875):
    """The random numbers of one pass of the chain: params [B,6], jitter [B,N,3], drop_u [B,N] (all on `device`)."""
    u = torch.rand((B, 6), device=device, generator=generator, dtype=F32)
    params = torch.empty((B, 6), device=device, dtype=F32)
    params[:, 0] = lo + (hi - lo) * u[:, 0]                          # PointcloudScale
    params[:, 1] = 2 * math.pi * u[:, 1]                             # PointcloudRotate
    params[:, 2:5] = translate_range * (2 * u[:, 2:5] - 1)           # PointcloudTranslate
    params[:, 5] = max_dropout_ratio * u[:, 5]                       # PointcloudRandomInputDropout
    jitter = torch.randn((B, N, 3), device=device, generator=generator, dtype=F32) * jitter_std
    drop_u = torch.rand((B, N), device=device, generator=generator, dtype=F32)
    return params, jitter, drop_u


def augment_clouds(pts, draws=None, generator=None, jitter_clip=0.05):
    """pts [B,N,3] fp32 CUDA -> augmented [B,N,3].  draws = (params [B,6], jitter [B,N,3], drop_u [B,N]) or None (drawn here)."""
    _lib.require_cuda(pts)
    pts = pts.float().contiguous()
    B, N, C = pts.shape
    if C != 3:
        raise NotImplementedError("augment_clouds handles xyz clouds (point_channels == 3)")
    params, jitter, drop_u = draws if draws is not None else draw(B, N, pts.device, generator)
    out = torch.empty_like(pts)
    _lib.call("vpf_augment_clouds", _lib.ptr(pts), _lib.ptr(params.float().contiguous()), _lib.ptr(jitter.float().contiguous()),
              _lib.ptr(drop_u.float().contiguous()), _lib.ptr(out), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_float(jitter_clip),
              _lib.stream_ptr())
    return out


class DeviceAugment:
    """trans_1 / trans_2 of datasets/data.py as a callable over a device batch."""

    def __init__(self, generator=None):
        self.generator = generator

    def __call__(self, pts):
        return augment_clouds(pts, generator=self.generator)
