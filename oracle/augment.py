"""numpy restatement of the reference's point-cloud augmentation chain -- TEST INFRASTRUCTURE ONLY.

trans_1 / trans_2 of datasets/data.py:16-36 (identical chains), applied per cloud by ShapeNetRender.__getitem__
(data.py:109-112): PointcloudNormalize (data_utils.py:206-221) -> PointcloudScale(0.5, 2) (:56-66) -> PointcloudRotate about
y (:69-98, angle_axis :6-34) -> PointcloudTranslate(0.5) (:156-171) -> PointcloudJitter(0.01, 0.05) (:141-153) ->
PointcloudRandomInputDropout(0.875) (:179-193).  The reference draws from numpy's / torch's global RNGs; here the draws
are explicit inputs, so the restatement can be pinned to the real classes (tests/make_golden_aug.py replays their draw
sequence) and the CUDA kernel can be checked bit-for-tolerance on the same draws.
"""
import numpy as np


def augment_cloud(pc, scaler, angle, trans, jitter, drop_ratio, drop_u):
    """pc [N,3] float32; scaler, angle, drop_ratio scalars; trans [3] in [-0.5, 0.5]; jitter [N,3] ~ N(0, 0.01) BEFORE the
    clamp; drop_u [N] uniform.  Returns the augmented float32 cloud."""
    pc = np.asarray(pc, np.float32).copy()
    centroid = np.mean(pc, axis=0)                       # PointcloudNormalize
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    pc = (pc / m).astype(np.float32)
    pc = pc * np.float32(scaler)                         # PointcloudScale
    c, s = np.float32(np.cos(angle)), np.float32(np.sin(angle))
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)   # angle_axis(angle, [0, 1, 0]).float()
    pc = (pc @ R.T).astype(np.float32)                   # PointcloudRotate
    diff = pc.max(axis=0) - pc.min(axis=0)               # PointcloudTranslate
    pc = (pc + (np.asarray(trans, np.float64) * diff).astype(np.float64)).astype(np.float32)
    pc = pc + np.clip(np.asarray(jitter, np.float32), -0.05, 0.05)   # PointcloudJitter
    drop = np.asarray(drop_u, np.float32) <= np.float32(drop_ratio)  # PointcloudRandomInputDropout
    if drop.any():
        pc[drop] = pc[0]
    return pc.astype(np.float32)


def augment_batch(pts, draws):
    """pts [B,N,3]; draws: dict of arrays scaler [B], angle [B], trans [B,3], jitter [B,N,3], drop_ratio [B], drop_u [B,N]."""
    return np.stack([augment_cloud(pts[b], draws["scaler"][b], draws["angle"][b], draws["trans"][b], draws["jitter"][b],
                                   draws["drop_ratio"][b], draws["drop_u"][b]) for b in range(pts.shape[0])])
