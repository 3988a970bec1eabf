"""Point-cloud tokenizer: drop-in for the reference's vipformer/model/pointcloud/utils.py.

Same names, argument order and return conventions as the reference functions
(divide_patches:6-38, fps:41-53, farthest_point_sample:56-85,
index_points:88-104, knn_point:107-119, square_distance:122-141); every one
runs a hand-written sm_100a kernel through the C ABI (include/vpf.h).

Two deliberate, documented refinements of the reference contract:
  * FPS start indices (reference: torch.randint from the global RNG,
    utils.py:71) can be passed explicitly (`start_idx=`) or drawn from a
    `generator=`; by default they are drawn with torch.randint on the input's
    device, exactly like the reference.
  * kNN output order (reference: torch.topk(sorted=False), unspecified) is
    fixed to (distance ascending, index ascending).  The reference's
    slot-axis slicing quirk (utils.py:36: only neighbour slots 0..2 get the
    centre subtracted) is reproduced under that order.
"""
import ctypes

import torch

from ... import _lib

_c_int = ctypes.c_int


def _prep(t, dtype=torch.float32):
    _lib.require_cuda(t)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _start(pts, start_idx, generator):
    B, N, _ = pts.shape
    if start_idx is None:
        # utils.py:71
        return torch.randint(0, N, (B,), dtype=torch.long, device=pts.device, generator=generator)
    start_idx = torch.as_tensor(start_idx, dtype=torch.long, device=pts.device).contiguous()
    if start_idx.shape != (B,):
        raise ValueError(f"start_idx must have shape ({B},), got {tuple(start_idx.shape)}")
    return start_idx


def farthest_point_sample(pts, npoint, *, start_idx=None, generator=None):
    """utils.py:56-85.  pts [B,N,C] -> LongTensor [B,npoint]."""
    pts = _prep(pts)
    B, N, C = pts.shape
    start = _start(pts, start_idx, generator)
    out = torch.empty((B, npoint), dtype=torch.long, device=pts.device)
    _lib.call("vpf_fps", _lib.ptr(pts), _c_int(B), _c_int(N), _c_int(C), _c_int(npoint),
              _lib.ptr(start), _lib.ptr(out), _lib.stream_ptr())
    return out


def index_points(points, idx):
    """utils.py:88-104.  points [B,N,C], idx [B,S] (or [B,S1,S2]) -> [B,S...,C]."""
    points = _prep(points)
    idx = _prep(idx, torch.long)
    B, N, C = points.shape
    flat = idx.reshape(B, -1)
    out = torch.empty((B, flat.shape[1], C), dtype=torch.float32, device=points.device)
    _lib.call("vpf_index_points", _lib.ptr(points), _c_int(B), _c_int(N), _c_int(C), _lib.ptr(flat),
              _c_int(flat.shape[1]), _lib.ptr(out), _lib.stream_ptr())
    return out.reshape(*idx.shape, C)


def fps(pts, number, *, start_idx=None, generator=None):
    """utils.py:41-53.  -> [B,number,C]."""
    return index_points(pts, farthest_point_sample(pts, number, start_idx=start_idx, generator=generator))


def square_distance(src, dst):
    """utils.py:122-141.  src [B,N,C], dst [B,M,C] -> [B,N,M] (expanded form, pinned arithmetic)."""
    src, dst = _prep(src), _prep(dst)
    B, S, Cs = src.shape
    _, N, Cd = dst.shape
    out = torch.empty((B, S, N), dtype=torch.float32, device=src.device)
    _lib.call("vpf_square_distance", _lib.ptr(src), _c_int(B), _c_int(S), _c_int(Cs), _lib.ptr(dst),
              _c_int(N), _c_int(Cd), _lib.ptr(out), _lib.stream_ptr())
    return out


def knn_point(nsample, xyz, new_xyz):
    """utils.py:107-119.  xyz [B,N,C], new_xyz [B,S,C] -> LongTensor [B,S,nsample]."""
    xyz, new_xyz = _prep(xyz), _prep(new_xyz)
    B, N, C = xyz.shape
    _, S, Cq = new_xyz.shape
    out = torch.empty((B, S, nsample), dtype=torch.long, device=xyz.device)
    _lib.call("vpf_knn_point", _c_int(nsample), _lib.ptr(xyz), _c_int(B), _c_int(N), _c_int(C),
              _lib.ptr(new_xyz), _c_int(S), _c_int(Cq), _lib.ptr(out), _lib.stream_ptr())
    return out


def divide_patches(points, num_groups, group_size, *, start_idx=None, generator=None, return_indices=False):
    """utils.py:6-38.  points [B,N,C] -> (neighbors [B,G,S,C], centers [B,G,C]).

    One FPS kernel + one fused kNN/gather kernel; no [B,G,N] distance matrix.
    """
    points = _prep(points)
    B, N, C = points.shape
    start = _start(points, start_idx, generator)
    dev = points.device
    neighbors = torch.empty((B, num_groups, group_size, C), dtype=torch.float32, device=dev)
    centers = torch.empty((B, num_groups, C), dtype=torch.float32, device=dev)
    fi = ki = None
    if return_indices:
        fi = torch.empty((B, num_groups), dtype=torch.long, device=dev)
        ki = torch.empty((B, num_groups, group_size), dtype=torch.long, device=dev)
    _lib.call("vpf_divide_patches", _lib.ptr(points), _c_int(B), _c_int(N), _c_int(C), _c_int(num_groups),
              _c_int(group_size), _lib.ptr(start), _lib.ptr(neighbors), _lib.ptr(centers), _lib.ptr(fi),
              _lib.ptr(ki), _lib.stream_ptr())
    if return_indices:
        return neighbors, centers, fi, ki
    return neighbors, centers
