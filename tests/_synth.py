"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

Shapes follow SURVEY.md section 8(d): clouds are randn, centred and scaled to
unit max-norm (mimics PointcloudNormalize, datasets/data_utils.py:206-221 of
the reference); the adversarial variants reproduce what
PointcloudRandomInputDropout (data_utils.py:174-190) does to real inputs --
a large fraction of points overwritten by point 0, i.e. exact duplicates.
"""
import numpy as np


def normalize_cloud(p):
    p = p - p.mean(axis=1, keepdims=True)
    m = np.sqrt((p ** 2).sum(-1)).max(axis=1, keepdims=True)[..., None]
    return (p / m).astype(np.float32)


def make_clouds(kind, B, N, seed):
    rng = np.random.default_rng(seed)
    if kind == "randn":
        return normalize_cloud(rng.standard_normal((B, N, 3)).astype(np.float32))
    if kind.startswith("dup"):  # dup50, dup875: fraction of points overwritten with point 0
        frac = {"dup50": 0.5, "dup875": 0.875}[kind]
        p = normalize_cloud(rng.standard_normal((B, N, 3)).astype(np.float32))
        for b in range(B):
            sel = rng.random(N) < frac
            p[b, sel] = p[b, 0]
        return p
    if kind == "grid":  # co-planar integer lattice: masses of exact distance ties
        side = int(np.ceil(np.sqrt(N)))
        g = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)[:N]
        p = np.zeros((B, N, 3), dtype=np.float32)
        for b in range(B):
            perm = rng.permutation(N)
            p[b, :, :2] = g[perm] * np.float32(0.125)
            p[b, :, 2] = np.float32(b)
        return p
    if kind == "allsame":
        return np.full((B, N, 3), 0.25, dtype=np.float32)
    raise ValueError(kind)


def make_start(B, N, seed):
    return np.random.default_rng(seed + 7919).integers(0, N, size=(B,), dtype=np.int64)


# (name, kind, B, N, G, S, seed) -- tokenizer golden / parity cases
TOKENIZER_CASES = [
    ("randn_1024_96", "randn", 4, 1024, 96, 32, 11),
    ("randn_2048_128", "randn", 3, 2048, 128, 32, 12),
    ("randn_2500_128", "randn", 2, 2500, 128, 32, 13),
    ("dup50_2048_128", "dup50", 2, 2048, 128, 32, 14),
    ("dup875_1024_96", "dup875", 2, 1024, 96, 32, 15),
    ("grid_1024_96", "grid", 2, 1024, 96, 32, 16),
    ("small_64_96", "randn", 2, 64, 96, 32, 17),     # N < G: FPS exhausts the cloud
    ("allsame_128_16", "allsame", 1, 128, 16, 32, 18),
    ("odd_333_40_8", "randn", 3, 333, 40, 8, 19),     # ragged N, non-default S
]
