"""CPU: pin the C oracle against golden vectors generated from the real reference."""
import os

import numpy as np
import pytest

import _synth
from _check import check_tokenizer_against_golden
from oracle import tokenizer as T


@pytest.mark.parametrize("case", _synth.TOKENIZER_CASES, ids=[c[0] for c in _synth.TOKENIZER_CASES])
def test_tokenizer_oracle_matches_reference(case, golden_dir):
    name, kind, B, N, G, S, seed = case
    g = np.load(os.path.join(golden_dir, f"tok_{name}.npz"))
    pts = _synth.make_clouds(kind, B, N, seed)
    assert np.float64(pts.astype(np.float64).sum()) == g["pts_sum"], "synthetic generator drifted from the fixture"
    start = _synth.make_start(B, N, seed)
    assert np.array_equal(start, g["start"])
    S = min(S, N)
    nb, ce, fi, ki = T.divide_patches(pts, G, S, start, return_indices=True)
    d = T.square_distance(ce, pts)
    frac = check_tokenizer_against_golden(g, fi, ce, ki, nb, lambda idx: np.take_along_axis(d, idx, 2))
    if kind == "randn":
        assert frac > 0.99  # ties are essentially absent on generic data: indices are bit-exact
        assert np.array_equal(np.sort(ki, -1), g["knn_set"])


def test_oracle_piecewise_functions_agree():
    pts = _synth.make_clouds("randn", 2, 512, 5)
    start = _synth.make_start(2, 512, 5)
    fi = T.farthest_point_sample(pts, 64, start)
    ce = T.index_points(pts, fi)
    assert np.array_equal(ce, T.fps(pts, 64, start))
    ki = T.knn_point(16, pts, ce)
    nb, ce2, fi2, ki2 = T.divide_patches(pts, 64, 16, start, return_indices=True)
    assert np.array_equal(fi, fi2) and np.array_equal(ce, ce2) and np.array_equal(ki, ki2)
    # slot quirk (reference utils.py:36): slots 0..2 centred, slots >= 3 absolute
    raw = np.take_along_axis(pts[:, None], ki[..., None].repeat(3, -1), 2)
    assert np.array_equal(nb[:, :, 3:], raw[:, :, 3:])
    assert np.array_equal(nb[:, :, :3], raw[:, :, :3] - ce[:, :, None])
    assert np.all(nb[:, :, 0] == 0)  # slot 0 is the centre itself


def test_fps_properties():
    pts = _synth.make_clouds("randn", 3, 256, 9)
    start = np.array([0, 100, 255], dtype=np.int64)
    fi = T.farthest_point_sample(pts, 256, start)
    assert np.array_equal(fi[:, 0], start)
    for b in range(3):  # sampling all N points of a duplicate-free cloud is a permutation
        assert sorted(fi[b].tolist()) == list(range(256))
    with pytest.raises(RuntimeError):
        T.farthest_point_sample(pts, 4, np.array([0, 0, 256], dtype=np.int64))


def test_torch_tokenizer_restatement_matches_c_oracle():
    """oracle/tokenizer_torch.py (the reference's ATen op sequence, used by bench.py's GPU-eager comparator) against the C
    oracle on generic clouds: FPS indices exact, neighbour SETS equal, centres exact; order equal where distances are untied."""
    import numpy as np
    import torch

    import _synth
    from oracle import tokenizer as T
    from oracle import tokenizer_torch as TT

    pts = _synth.make_clouds("randn", 3, 1024, 5)
    start = _synth.make_start(3, 1024, 5)
    nb, ce, fi, ki = T.divide_patches(pts, 96, 32, start, return_indices=True)
    tnb, tce = TT.divide_patches(torch.from_numpy(pts), 96, 32, torch.from_numpy(start))
    assert np.array_equal(tce.numpy(), ce)
    assert np.array_equal(TT.farthest_point_sample(torch.from_numpy(pts), 96, torch.from_numpy(start)).numpy(), fi)
    same = (tnb.numpy() == nb).all(-1)
    assert same.mean() > 0.999       # identical up to fp32 ties between the expanded-form bmm and the pinned C arithmetic


def test_augment_oracle_matches_reference_golden(golden_dir):
    """oracle/augment.py against vectors from the reference's own data_utils classes (tests/make_golden_aug.py)."""
    import numpy as np

    from oracle import augment as A

    g = np.load(os.path.join(golden_dir, "aug_trans1.npz"))
    out = A.augment_batch(g["raw"], {k: g[k] for k in ("scaler", "angle", "trans", "jitter", "drop_ratio", "drop_u")})
    assert np.abs(out - g["out"]).max() < 2e-6
