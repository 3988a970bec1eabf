"""Comparison helpers shared by the CPU and GPU parity tests."""
import numpy as np


def check_tokenizer_against_golden(g, fps_idx, centers, knn_idx, neighbors, dist_of):
    """Compare a tokenizer result with a golden fixture made from the reference.

    The reference's kNN order under exact distance ties is arbitrary
    (torch.topk); the stated rule here is lowest-index-first.  So:
      * FPS indices and centres: bit-exact, always.
      * the vector of selected distances (ascending): bit-exact, always.
      * kNN indices and gathered neighbours: bit-exact at every slot whose
        distance is strictly separated from its row neighbours (and, for the
        last slot, from the first excluded point).
    `dist_of(idx)` returns the pinned expanded-form distance for selected idx.
    """
    assert np.array_equal(fps_idx, g["fps_idx"]), "FPS indices differ from the reference"
    assert np.array_equal(centers, g["centers"]), "centres differ from the reference"
    kd = dist_of(knn_idx)
    assert np.array_equal(kd, g["knn_dist"]), "selected distances differ from the reference"
    gd = g["knn_dist"]
    S = gd.shape[-1]
    nxt = np.concatenate([gd[..., 1:], g["next_dist"][..., None]], -1)
    prv = np.concatenate([np.full_like(gd[..., :1], -np.inf), gd[..., :-1]], -1)
    unique = (gd < nxt) & (gd > prv)
    if g["next_dist"].shape[-1:] == () or S == 0:
        pass
    assert np.array_equal(knn_idx[unique], g["knn_idx"][unique]), "kNN indices differ at untied slots"
    assert np.array_equal(neighbors[unique], g["neighbors"][unique]), "neighbours differ at untied slots"
    return float(unique.mean())
