"""ctypes binding of oracle/tokenizer_oracle.c -- TEST INFRASTRUCTURE ONLY.

numpy in / numpy out; mirrors the reference signatures in
vipformer/model/pointcloud/utils.py (fps:41-53, farthest_point_sample:56-85,
index_points:88-104, knn_point:107-119, square_distance:122-141,
divide_patches:6-38) with the FPS start index made an explicit argument.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvpf_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "tokenizer_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for name in ("vpf_oracle_fps", "vpf_oracle_index_points", "vpf_oracle_square_distance",
                     "vpf_oracle_knn", "vpf_oracle_divide_patches"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with oracle error {rc}")


def farthest_point_sample(pts, npoint, start_idx):
    pts = _f32(pts)
    B, N, C = pts.shape
    start_idx = _i64(start_idx)
    out = np.empty((B, npoint), dtype=np.int64)
    _check(lib().vpf_oracle_fps(_p(pts), B, N, C, int(npoint), _p(start_idx), _p(out)), "fps")
    return out


def index_points(points, idx):
    points = _f32(points)
    idx = _i64(idx)
    B, N, C = points.shape
    S = idx.shape[1]
    out = np.empty((B, S, C), dtype=np.float32)
    _check(lib().vpf_oracle_index_points(_p(points), B, N, C, _p(idx), S, _p(out)), "index_points")
    return out


def fps(pts, number, start_idx):
    return index_points(pts, farthest_point_sample(pts, number, start_idx))


def square_distance(src, dst):
    src, dst = _f32(src), _f32(dst)
    B, S, Cq = src.shape
    _, N, C = dst.shape
    out = np.empty((B, S, N), dtype=np.float32)
    _check(lib().vpf_oracle_square_distance(_p(src), B, S, Cq, _p(dst), N, C, _p(out)), "square_distance")
    return out


def knn_point(nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, C = xyz.shape
    _, S, Cq = new_xyz.shape
    out = np.empty((B, S, nsample), dtype=np.int64)
    _check(lib().vpf_oracle_knn(int(nsample), _p(xyz), B, N, C, _p(new_xyz), S, Cq, _p(out)), "knn_point")
    return out


def divide_patches(points, num_groups, group_size, start_idx, return_indices=False):
    points = _f32(points)
    B, N, C = points.shape
    start_idx = _i64(start_idx)
    nb = np.empty((B, num_groups, group_size, C), dtype=np.float32)
    ce = np.empty((B, num_groups, C), dtype=np.float32)
    fi = np.empty((B, num_groups), dtype=np.int64)
    ki = np.empty((B, num_groups, group_size), dtype=np.int64)
    _check(lib().vpf_oracle_divide_patches(_p(points), B, N, C, int(num_groups), int(group_size),
                                           _p(start_idx), _p(nb), _p(ce), _p(fi), _p(ki)), "divide_patches")
    if return_indices:
        return nb, ce, fi, ki
    return nb, ce
