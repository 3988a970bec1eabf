"""Generate tests/golden/*.npz from the REAL reference (run in the build container).

    python tests/make_golden.py [tokenizer] [model] [loss]

The reference (/root/reference) cannot travel to the GPU box, so its outputs on
seeded synthetic inputs are committed as small fixtures.  Pinning applied to
the reference (both are legal refinements of its contract):
  * FPS start index: `torch.randint` in the reference's namespace is replaced
    by a function returning the fixture's explicit start indices
    (vipformer/model/pointcloud/utils.py:71).
  * kNN order: `torch.topk(..., sorted=False)` (utils.py:117) is run with
    sorted=True.  The un-patched index SETS are stored too.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _refshim  # noqa: E402
import _synth  # noqa: E402

GOLD = os.path.join(HERE, "golden")


class _PinnedRandint:
    def __init__(self, start):
        self.start = torch.as_tensor(start, dtype=torch.long)

    def __call__(self, low, high, size, **kw):
        assert tuple(size) == tuple(self.start.shape)
        return self.start.clone()


def tokenizer_reference(pts, G, S, start):
    """Run the reference divide_patches with pinned start + sorted top-k."""
    _refshim.load()
    import vipformer.model.pointcloud.utils as U

    class _T:  # torch proxy whose randint is pinned; everything else forwards
        def __getattr__(self, k):
            return getattr(torch, k)

    proxy = _T()
    proxy.randint = _PinnedRandint(start)
    orig_torch, orig_knn = U.torch, U.knn_point

    def knn_sorted(nsample, xyz, new_xyz):
        d = U.square_distance(new_xyz, xyz)
        return torch.topk(d, nsample, dim=-1, largest=False, sorted=True)[1]

    t = torch.from_numpy(pts)
    try:
        U.torch = proxy
        fps_idx = U.farthest_point_sample(t, G)
        U.knn_point = knn_sorted
        nb, ce = U.divide_patches(t.clone(), G, S)
        knn_sorted_idx = knn_sorted(S, t[:, :, :3], ce[:, :, :3])
        U.knn_point = orig_knn
        knn_unsorted_idx = U.knn_point(S, t[:, :, :3], ce[:, :, :3])
        dist = U.square_distance(ce[:, :, :3], t[:, :, :3])
    finally:
        U.torch, U.knn_point = orig_torch, orig_knn
    kd = torch.gather(dist, 2, knn_sorted_idx)
    # the (S+1)-th smallest distance: lets a checker tell a genuine tie at the cut
    kth1 = torch.topk(dist, min(S + 1, dist.shape[-1]), dim=-1, largest=False, sorted=True)[0][..., -1]
    return dict(fps_idx=fps_idx.numpy(), neighbors=nb.numpy(), centers=ce.numpy(),
                knn_idx=knn_sorted_idx.numpy(), knn_set=np.sort(knn_unsorted_idx.numpy(), -1),
                knn_dist=kd.numpy(), next_dist=kth1.numpy())


def gen_tokenizer():
    for name, kind, B, N, G, S, seed in _synth.TOKENIZER_CASES:
        pts = _synth.make_clouds(kind, B, N, seed)
        start = _synth.make_start(B, N, seed)
        out = tokenizer_reference(pts, G, min(S, N), start)
        # inputs are regenerated from the seed by the tests; store a checksum only
        np.savez_compressed(os.path.join(GOLD, f"tok_{name}.npz"), start=start,
                            pts_sum=np.float64(pts.astype(np.float64).sum()), **out)
        print("tokenizer", name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1:] or ["tokenizer", "model", "loss"]
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "tokenizer" in what:
        gen_tokenizer()
    if "model" in what or "loss" in what:
        import make_golden_model
        make_golden_model.main(what)
