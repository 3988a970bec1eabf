"""Golden vectors for the augmentation chain from the REAL reference classes (build container only).

Runs trans_1 of datasets/data.py:16-25, rebuilt from the reference's own datasets/data_utils.py classes (data.py itself
imports h5py / PIL, absent here), under seeded numpy / torch RNGs, then REPLAYS the same draw sequence to record the draws:
Scale: uniform(0,1), uniform(lo,hi) | Rotate: uniform(0,1), uniform() | Translate: uniform(0,1), uniform(-r,r,3) |
Jitter: uniform(0,1), torch normal_ | Dropout: uniform(0,1), random(), random(N).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import augment as A  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_data_utils", "/root/reference/datasets/data_utils.py")
D = importlib.util.module_from_spec(spec)
spec.loader.exec_module(D)


def chain():
    return [D.PointcloudToTensor(), D.PointcloudNormalize(), D.PointcloudScale(lo=0.5, hi=2, p=1), D.PointcloudRotate(),
            D.PointcloudTranslate(0.5, p=1), D.PointcloudJitter(p=1), D.PointcloudRandomInputDropout(p=1)]


def main():
    B, N = 6, 1024
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal((B, N, 3)) * np.array([1.0, 0.5, 2.0]) + np.array([3.0, -1.0, 0.5])).astype(np.float32)
    outs, draws = [], dict(scaler=[], angle=[], trans=[], jitter=[], drop_ratio=[], drop_u=[])
    for b in range(B):
        np.random.seed(100 + b)
        torch.manual_seed(200 + b)
        x = raw[b].copy()
        for t in chain():
            x = t(x)
        outs.append(x.numpy())
        np.random.seed(100 + b)
        torch.manual_seed(200 + b)
        np.random.uniform(0, 1); draws["scaler"].append(np.random.uniform(0.5, 2))
        np.random.uniform(0, 1); draws["angle"].append(np.random.uniform() * 2 * np.pi)
        np.random.uniform(0, 1); draws["trans"].append(np.random.uniform(-0.5, 0.5, size=(3)))
        np.random.uniform(0, 1); draws["jitter"].append(torch.empty(N, 3).normal_(mean=0.0, std=0.01).numpy())
        np.random.uniform(0, 1); ratio = np.random.random() * 0.875; u = np.random.random((N))
        assert np.array_equal(u <= ratio, u.astype(np.float32) <= np.float32(ratio)), "float32 compare would differ"
        draws["drop_ratio"].append(ratio); draws["drop_u"].append(u)
    out = np.stack(outs)
    draws = {k: np.asarray(v) for k, v in draws.items()}
    mine = A.augment_batch(raw, draws)
    err = np.abs(mine - out).max()
    print("oracle vs real reference max abs err", err, "dropped fraction", float((draws["drop_u"] <= draws["drop_ratio"][:, None]).mean()))
    assert err < 2e-6
    np.savez_compressed(os.path.join(HERE, "golden", "aug_trans1.npz"), raw=raw, out=out, **draws)


if __name__ == "__main__":
    main()
