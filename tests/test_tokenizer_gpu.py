"""GPU: CUDA tokenizer vs the C oracle (bit-exact) and vs the reference goldens."""
import os

import numpy as np
import pytest
import torch

import _synth
from _check import check_tokenizer_against_golden
from oracle import tokenizer as T

pytestmark = pytest.mark.gpu


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", _synth.TOKENIZER_CASES, ids=[c[0] for c in _synth.TOKENIZER_CASES])
def test_divide_patches_bit_exact_vs_oracle_and_golden(case, golden_dir):
    from vipformer_b200.model.pointcloud import utils as U

    name, kind, B, N, G, S, seed = case
    S = min(S, N)
    pts = _synth.make_clouds(kind, B, N, seed)
    start = _synth.make_start(B, N, seed)
    nb, ce, fi, ki = U.divide_patches(_gpu(pts), G, S, start_idx=_gpu(start), return_indices=True)
    nb, ce, fi, ki = (t.cpu().numpy() for t in (nb, ce, fi, ki))
    onb, oce, ofi, oki = T.divide_patches(pts, G, S, start, return_indices=True)
    assert np.array_equal(fi, ofi)
    assert np.array_equal(ce, oce)
    assert np.array_equal(ki, oki)          # bit-exact incl. ties (lowest index first)
    assert np.array_equal(nb, onb)
    g = np.load(os.path.join(golden_dir, f"tok_{name}.npz"))
    d = U.square_distance(_gpu(ce), _gpu(pts)).cpu().numpy()
    assert np.array_equal(d, T.square_distance(ce, pts))
    check_tokenizer_against_golden(g, fi, ce, ki, nb, lambda idx: np.take_along_axis(d, idx, 2))


def test_piecewise_api_matches_oracle():
    from vipformer_b200.preproc import farthest_point_sample, fps, index_points, knn_point

    pts = _synth.make_clouds("dup50", 5, 1500, 3)
    start = _synth.make_start(5, 1500, 3)
    fi = farthest_point_sample(_gpu(pts), 100, start_idx=_gpu(start))
    assert fi.dtype == torch.long
    assert np.array_equal(fi.cpu().numpy(), T.farthest_point_sample(pts, 100, start))
    ce = fps(_gpu(pts), 100, start_idx=_gpu(start))
    assert np.array_equal(ce.cpu().numpy(), T.fps(pts, 100, start))
    assert np.array_equal(index_points(_gpu(pts), fi).cpu().numpy(), ce.cpu().numpy())
    ki = knn_point(20, _gpu(pts), ce)
    assert np.array_equal(ki.cpu().numpy(), T.knn_point(20, pts, ce.cpu().numpy()))


@pytest.mark.parametrize("B,N,G", [(64, 2048, 128), (96, 1024, 96)])
def test_full_size_properties(B, N, G):
    """BASELINE-size batch: size-independent properties + oracle on a sub-sample of clouds."""
    from vipformer_b200.preproc import divide_patches

    pts = _synth.make_clouds("randn", B, N, 77)
    start = _synth.make_start(B, N, 77)
    nb, ce, fi, ki = divide_patches(_gpu(pts), G, 32, start_idx=_gpu(start), return_indices=True)
    nb, ce, fi, ki = (t.cpu().numpy() for t in (nb, ce, fi, ki))
    assert np.array_equal(fi[:, 0], start)
    assert all(len(set(r.tolist())) == G for r in fi)            # FPS never repeats on distinct points
    assert np.array_equal(ki[:, :, 0], fi)                       # nearest neighbour of a centre is itself
    assert all(len(set(r.tolist())) == 32 for r in ki.reshape(-1, 32)[::37])
    assert np.all(nb[:, :, 0] == 0)
    raw = np.take_along_axis(pts[:, None], ki[..., None].repeat(3, -1), 2)
    assert np.array_equal(nb[:, :, 3:], raw[:, :, 3:])
    sub = slice(0, B, max(1, B // 6))
    onb, oce, ofi, oki = T.divide_patches(pts[sub], G, 32, start[sub], return_indices=True)
    assert np.array_equal(fi[sub], ofi) and np.array_equal(ki[sub], oki) and np.array_equal(nb[sub], onb)


def test_channels_gt3_and_host_entry():
    import ctypes
    from vipformer_b200 import _lib
    from vipformer_b200.preproc import divide_patches

    rng = np.random.default_rng(0)
    pts = rng.standard_normal((3, 700, 6)).astype(np.float32)
    start = _synth.make_start(3, 700, 1)
    nb, ce = divide_patches(_gpu(pts), 50, 16, start_idx=_gpu(start))
    onb, oce = T.divide_patches(pts, 50, 16, start)
    assert np.array_equal(nb.cpu().numpy(), onb) and np.array_equal(ce.cpu().numpy(), oce)
    # host-buffer entry (what bench.py's e2e leg times)
    p3 = np.ascontiguousarray(pts[:, :, :3])
    nbytes = _lib.size_query("vpf_divide_patches_host_workspace_bytes", 3, 700, 3, 50, 16)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    hnb = np.empty((3, 50, 16, 3), np.float32)
    hce = np.empty((3, 50, 3), np.float32)
    _lib.call("vpf_divide_patches_host", p3.ctypes.data_as(ctypes.c_void_p), 3, 700, 3, 50, 16,
              start.ctypes.data_as(ctypes.c_void_p), hnb.ctypes.data_as(ctypes.c_void_p),
              hce.ctypes.data_as(ctypes.c_void_p), _lib.ptr(ws), ctypes.c_size_t(nbytes), _lib.stream_ptr())
    onb3, oce3 = T.divide_patches(p3, 50, 16, start)
    assert np.array_equal(hnb, onb3) and np.array_equal(hce, oce3)


def test_errors():
    from vipformer_b200 import _lib
    from vipformer_b200.preproc import divide_patches, knn_point

    x = torch.randn(2, 100, 3, device="cuda")
    with pytest.raises(_lib.VpfError):
        divide_patches(x, 8, 64)          # group_size > 32 unsupported
    with pytest.raises(_lib.VpfError):
        knn_point(16, x[:, :8], x[:, :4])  # nsample > N
    with pytest.raises(ValueError):
        divide_patches(x, 8, 4, start_idx=torch.zeros(3, dtype=torch.long))
