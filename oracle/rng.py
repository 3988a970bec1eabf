"""numpy restatement of the product's counter-based dropout masks -- TEST INFRASTRUCTURE ONLY.

The reference draws nn.Dropout masks from torch's global RNG (partseg.py:81 attention probabilities,
partseg.py:208-213 Residual); a from-scratch kernel cannot reproduce that stream, so dropout-ON parity is
defined as: the oracle, given the SAME keep-masks the kernels use, must produce the same outputs.  The masks are
a pure function of (seed, op_id, element index) -- vipformer_b200/csrc/rng.cuh and attention*.cu -- restated here.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def mix32(x):
    """lowbias32 finaliser (rng.cuh mix32) on uint32 arrays."""
    x = np.asarray(x, dtype=np.uint64) & _M
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M
    x ^= x >> np.uint64(16)
    return x


def make_key(seed, op_id):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & 0xFFFFFFFF, seed >> 32
    inner = int(mix32((hi + 0x9E3779B9 * ((op_id + 1) & 0xFFFFFFFF)) & 0xFFFFFFFF))
    return int(mix32(lo ^ inner))


def threshold8(p):
    """rng.cuh threshold8: drop iff byte < round(256 p) (clamped to 255)."""
    if not p > 0.0:
        return 0
    return min(255, int(float(np.float32(p)) * 256.0 + 0.5))


def residual_keep(seed, op_id, p, rows, cols):
    """Residual / Dropout mask of a [rows, cols] tensor (gemm.cu residual epilogue, dropout_grad; rng.cuh keep8): element
    index e = row * cols + col keeps iff byte (e & 3) of hash(key, e >> 2) >= round(256 p); returns the float32 mask
    already scaled by 256 / (256 - round(256 p))  (p = 0.5 exact, p = 0.1 -> 26/256)."""
    thr = threshold8(p)
    if thr == 0:
        return np.ones((rows, cols), np.float32)
    key = np.uint64(make_key(seed, op_id))
    idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :]) & _M
    h = mix32((((idx >> np.uint64(2)) * np.uint64(0x9E3779B1)) & _M) ^ key)
    byte = (h >> ((idx & np.uint64(3)) * np.uint64(8))) & np.uint64(0xFF)
    keep = byte >= np.uint64(thr)
    return keep.astype(np.float32) * np.float32(256.0 / (256.0 - thr))


def attention_keep(seed, op_id, p, BH, Lq, Lk):
    """Attention-probability dropout mask [BH, Lq, Lk] (attention.cu / attention_tc.cu): one 32-bit hash covers 4
    adjacent keys of one query row, 8 bits each; effective p = round(256 p) / 256.  Scaled by 256 / (256 - thr)."""
    thr = int(float(np.float32(p)) * 256.0 + 0.5) if p > 0.0 else 0
    if thr == 0:
        return np.ones((BH, Lq, Lk), np.float32)
    key = np.uint64(make_key(seed, op_id))
    kq = (Lk + 3) >> 2
    bh = np.arange(BH, dtype=np.uint64)[:, None, None]
    i = np.arange(Lq, dtype=np.uint64)[None, :, None]
    j = np.arange(Lk, dtype=np.uint64)[None, None, :]
    idx = ((bh * np.uint64(Lq) + i) * np.uint64(kq) + (j >> np.uint64(2))) & _M
    h = mix32(((idx * np.uint64(0x9E3779B1)) & _M) ^ key)
    byte = (h >> ((j & np.uint64(3)) * np.uint64(8))) & np.uint64(0xFF)
    keep = byte >= np.uint64(thr)
    return keep.astype(np.float32) * np.float32(256.0 / (256.0 - thr))


def draw_indices(seed, op_id, n, N):
    """vpf_draw_indices (augment.cu): n indices in [0, N) from (seed, op_id): (hash * N) >> 32.  int64 array."""
    key = np.uint64(make_key(seed, op_id))
    i = np.arange(n, dtype=np.uint64)
    h = mix32(((i * np.uint64(0x9E3779B1)) & _M) ^ key)
    return ((h * np.uint64(N)) >> np.uint64(32)).astype(np.int64)


def droppath_scales(seed, op_id, p, B):
    """vpf_droppath_scales (pool.cu, rng.cuh keep_sample): float32 [B], 1 / (1 - p) for kept samples, 0 for dropped ones;
    sample b is dropped iff hash(key, b) < p * 2^32."""
    if not p > 0.0:
        return np.ones(B, np.float32)
    t = float(np.float32(p)) * 4294967296.0
    thr = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
    key = np.uint64(make_key(seed, op_id))
    b = np.arange(B, dtype=np.uint64)
    h = mix32(((b * np.uint64(0x9E3779B1)) & _M) ^ key)
    keep = h >= np.uint64(thr)
    return keep.astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p)))   # fp32 arithmetic, as the kernel
