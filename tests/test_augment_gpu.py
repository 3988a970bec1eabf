"""GPU: device-side augmentation against golden vectors from the REAL reference classes and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _draws(g, dev="cuda"):
    params = np.concatenate([g["scaler"][:, None], g["angle"][:, None], g["trans"], g["drop_ratio"][:, None]], 1).astype(np.float32)
    return (torch.from_numpy(params).to(dev), torch.from_numpy(g["jitter"].astype(np.float32)).to(dev),
            torch.from_numpy(g["drop_u"].astype(np.float32)).to(dev))


def test_augment_matches_reference_golden(golden_dir):
    from vipformer_b200.datasets import augment_clouds

    g = np.load(os.path.join(golden_dir, "aug_trans1.npz"))
    out = augment_clouds(torch.from_numpy(g["raw"]).cuda(), draws=_draws(g)).cpu().numpy()
    # fp32 arithmetic in a different summation order (centroid) than numpy's pairwise mean: 1e-5 absolute on O(1) coordinates
    assert np.abs(out - g["out"]).max() < 1e-5, np.abs(out - g["out"]).max()
    # the duplicates are EXACT copies of point 0 (what the tokenizer's tie-break rules exist for)
    drop = g["drop_u"].astype(np.float32) <= g["drop_ratio"].astype(np.float32)[:, None]
    for b in range(out.shape[0]):
        assert np.array_equal(out[b][drop[b]], np.broadcast_to(out[b, 0], (int(drop[b].sum()), 3)))


def test_augment_matches_oracle_at_pretraining_size():
    from oracle import augment as A
    from vipformer_b200.datasets import augment_clouds, draw

    B, N = 16, 2048
    gen = torch.Generator(device="cuda").manual_seed(3)
    raw = torch.randn((B, N, 3), device="cuda", generator=gen) * torch.tensor([1.0, 3.0, 0.3], device="cuda") + 2.0
    params, jitter, drop_u = draw(B, N, raw.device, gen)
    out = augment_clouds(raw, draws=(params, jitter, drop_u)).cpu().numpy()
    p = params.cpu().numpy().astype(np.float64)
    ref = A.augment_batch(raw.cpu().numpy(), dict(scaler=p[:, 0], angle=p[:, 1], trans=p[:, 2:5], jitter=jitter.cpu().numpy(),
                                                  drop_ratio=p[:, 5], drop_u=drop_u.cpu().numpy()))
    assert np.abs(out - ref).max() < 2e-5
    # properties of the chain: a second pass draws different views; the un-jittered cloud radius is the scale factor
    out2 = augment_clouds(raw, generator=gen)
    assert not torch.equal(out2.cpu(), torch.from_numpy(out))


def test_augmented_views_feed_the_tokenizer():
    """Two augmented views -> divide_patches: duplicates onto point 0 must not break FPS / kNN (bit-exact vs the C oracle)."""
    from oracle import tokenizer as T
    from vipformer_b200.datasets import DeviceAugment
    from vipformer_b200.preproc import divide_patches

    gen = torch.Generator(device="cuda").manual_seed(9)
    raw = torch.randn((4, 1024, 3), device="cuda", generator=gen)
    view = DeviceAugment(gen)(raw)
    start = torch.randint(0, 1024, (4,), device="cuda", generator=gen)
    nb, ce, fi, ki = divide_patches(view, 96, 32, start_idx=start, return_indices=True)
    onb, oce, ofi, oki = T.divide_patches(view.cpu().numpy(), 96, 32, start.cpu().numpy(), return_indices=True)
    assert np.array_equal(fi.cpu().numpy(), ofi) and np.array_equal(ki.cpu().numpy(), oki)
    assert np.array_equal(nb.cpu().numpy(), onb) and np.array_equal(ce.cpu().numpy(), oce)


@pytest.mark.parametrize("n,N", [(512, 2048), (7, 1), (1000, 1024), (33, 333)])
def test_draw_indices_bit_exact_vs_oracle(n, N):
    """FPS start points of the engine step (utils.py:71 draws them with torch.randint): integer work, bit-exact."""
    import oracle.rng as R
    from vipformer_b200 import ops

    state = torch.tensor([3, 0x1234567890ABCDEF - (1 << 63)], dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    ops.draw_indices(state, 0xF9500000, N, out)
    ref = R.draw_indices(int(state[1].item()) & 0xFFFFFFFFFFFFFFFF, 0xF9500000, n, N)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert int(out.min()) >= 0 and int(out.max()) < N
    if n >= 512:      # roughly uniform: every quarter of the range is hit
        assert len(set((out.cpu().numpy() * 4 // N).tolist())) == 4
