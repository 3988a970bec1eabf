"""Golden vectors for the floating-point half of the path, from the REAL reference (build container only).

For each case in _synth.MODEL_CASES: build the reference CrossFormer_pc_mp / CrossFormer_img_mp (seeded), check that
the product mirror builds the IDENTICAL state_dict (keys, shapes, values) under the same seed, perturb the weights
deterministically, run train-mode forward (dropout 0, batch-stat BatchNorm), the pretrain.py:189-207 loss
composition (NT-Xent restated: lightly is not installable here) and backward.  Stored: features, losses, updated
running statistics, per-parameter gradient norms and a few full gradients.
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
from make_golden import _PinnedRandint  # noqa: E402
from oracle import model_ref as M  # noqa: E402

GOLD = os.path.join(HERE, "golden")
FULL_GRADS = ["encoder.cross_attn_1.0.module.q_norm.weight", "group2emb.first_conv.0.weight", "latent_head.5.weight",
              "position_emb.0.weight", "input_adapter.point_mlp.1.bias", "encoder.sa_layers.0.1.module.1.bias"]


def run_case(name, cfg):
    _refshim.load()
    import vipformer.model.pointcloud.utils as U

    ref_pc, ref_im = _synth.build_models(cfg, pkg="vipformer")
    my_pc, my_im = _synth.build_models(cfg, pkg="vipformer_b200")
    for r, m in ((ref_pc, my_pc), (ref_im, my_im)):
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()), "state_dict keys differ from the reference"
        for k in rs:
            assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), f"state_dict value differs at {k}"
    sd_pc = _synth.perturb_state_dict(ref_pc.state_dict(), cfg["seed"] + 10)
    sd_im = _synth.perturb_state_dict(ref_im.state_dict(), cfg["seed"] + 11)
    ref_pc.load_state_dict(sd_pc)
    ref_im.load_state_dict(sd_im)
    ref_pc.train()
    ref_im.train()
    pts, start, imgs = _synth.model_inputs(cfg)

    class _T:
        def __getattr__(self, k):
            return getattr(torch, k)

    proxy = _T()
    proxy.randint = _PinnedRandint(start)
    orig_torch, orig_knn = U.torch, U.knn_point

    def knn_sorted(nsample, xyz, new_xyz):
        d = U.square_distance(new_xyz, xyz)
        return torch.topk(d, nsample, dim=-1, largest=False, sorted=True)[1]

    try:
        U.torch, U.knn_point = proxy, knn_sorted
        pc_feats, pc_back = ref_pc(pts)
    finally:
        U.torch, U.knn_point = orig_torch, orig_knn
    im_feats, im_back = ref_im(imgs)
    total, imid, cmid = M.pretrain_loss(pc_feats, im_feats)
    total.backward()
    out = dict(pc_feats=pc_feats.detach().numpy(), pc_backbone=pc_back.detach().numpy(),
               img_feats=im_feats.detach().numpy(), img_backbone=im_back.detach().numpy(),
               loss=np.array([total.item(), imid.item(), cmid.item()]))
    for tag, model in (("pc", ref_pc), ("img", ref_im)):
        names, norms = [], []
        for k, p in model.named_parameters():
            names.append(k)
            norms.append(p.grad.double().norm().item())
            if k in FULL_GRADS or k == "position_emb" or k == "patch2emb.1.bias":
                out[f"{tag}_grad::{k}"] = p.grad.numpy()
        out[f"{tag}_grad_names"] = np.array(names)
        out[f"{tag}_grad_norms"] = np.array(norms)
        for k, v in model.state_dict().items():
            if "running_" in k:
                out[f"{tag}_buf::{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **out)
    print("model", name, "loss", out["loss"], "pc_feats", out["pc_feats"].shape)


def run_ft_case(name, cfg):
    """CrossFormer_pc_mp_ft + CrossEntropyLoss(label_smoothing=0.2) (partseg.py:553-605, ft_cls.py:145,176) from the REAL
    reference: logits, loss, per-parameter gradient norms, a few full gradients, updated running statistics."""
    _refshim.load()
    import vipformer.model.pointcloud.utils as U

    ref = _synth.build_ft_model(cfg, pkg="vipformer")
    mine = _synth.build_ft_model(cfg, pkg="vipformer_b200")
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), "fine-tune state_dict keys differ from the reference"
    for k in rs:
        assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), f"state_dict value differs at {k}"
    sd = _synth.perturb_state_dict(ref.state_dict(), cfg["seed"] + 10)
    ref.load_state_dict(sd)
    ref.train()
    pts, start, labels = _synth.ft_inputs(cfg)

    class _T:
        def __getattr__(self, k):
            return getattr(torch, k)

    proxy = _T()
    proxy.randint = _PinnedRandint(start)
    orig_torch, orig_knn = U.torch, U.knn_point

    def knn_sorted(nsample, xyz, new_xyz):
        d = U.square_distance(new_xyz, xyz)
        return torch.topk(d, nsample, dim=-1, largest=False, sorted=True)[1]

    # partseg.py imported divide_patches by name: patch the functions it resolves at call time (module globals of utils)
    try:
        U.torch, U.knn_point = proxy, knn_sorted
        logits = ref(pts)
    finally:
        U.torch, U.knn_point = orig_torch, orig_knn
    loss = torch.nn.CrossEntropyLoss(label_smoothing=0.2)(logits, labels)
    loss.backward()
    out = dict(logits=logits.detach().numpy(), loss=np.array([loss.item()]))
    names, norms = [], []
    for k, p in ref.named_parameters():
        if p.grad is None:        # latent_head: unused by the fine-tune forward (find_unused_parameters=True, ft_cls.py:86)
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        if k in FULL_GRADS or k.startswith("finetune_head.8") or k == "finetune_head.2.weight":
            out[f"grad::{k}"] = p.grad.numpy()
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms)
    for k, v in ref.state_dict().items():
        if "running_" in k:
            out[f"buf::{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **out)
    print("ft model", name, "loss", out["loss"], "logits", out["logits"].shape, "params with grad", len(names))


def run_seg_case(name, cfg):
    """CrossFormer_partseg + CrossEntropyLoss(label_smoothing=0.2) over every point (partseg.py:345-470, ft_partseg.py:128,
    158-160) from the REAL reference (max_dpr = 0, pinned FPS start, sorted kNN): logits, loss, gradient norms, a few full
    gradients, updated running statistics."""
    _refshim.load()
    import vipformer.model.pointcloud.utils as U

    ref = _synth.build_seg_model(cfg, pkg="vipformer")
    mine = _synth.build_seg_model(cfg, pkg="vipformer_b200")
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), "part-segmentation state_dict keys differ from the reference"
    for k in rs:
        assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), f"state_dict value differs at {k}"
    ref.load_state_dict(_synth.perturb_state_dict(ref.state_dict(), cfg["seed"] + 10))
    ref.train()
    ref.dp1.p = 0.0          # nn.Dropout(0.5) draws from torch's RNG: off for the fixture (masks are tested on the GPU side)
    if cfg.get("max_dpr", 0.0) > 0.0:
        # DropPath: pin every instance's per-sample draw to oracle/rng.py droppath_scales(DPR_SEED, site id, rate)
        from oracle import rng as R
        ob, rates = _synth.seg_op_bases(cfg), _synth.seg_drop_path(cfg)
        for i, layer in enumerate(ref.encoder.sa_layers):
            key = f"seg.encoder.sa_layers.{i}"
            for which in (1, 2):
                dp = layer[which - 1].drop_path
                if key in rates:
                    assert type(dp).__name__ == "_DropPath" and abs(dp.p - rates[key]) < 1e-7
                    dp.scales = torch.from_numpy(R.droppath_scales(_synth.DPR_SEED, ob[key] + 2 + which, rates[key], cfg["b"]))
    pts, start, onehot, labels = _synth.seg_inputs(cfg)

    class _T:
        def __getattr__(self, k):
            return getattr(torch, k)

    proxy = _T()
    proxy.randint = _PinnedRandint(start)
    orig_torch, orig_knn = U.torch, U.knn_point

    def knn_sorted(nsample, xyz, new_xyz):
        d = U.square_distance(new_xyz, xyz)
        return torch.topk(d, nsample, dim=-1, largest=False, sorted=True)[1]

    try:
        U.torch, U.knn_point = proxy, knn_sorted
        logits = ref(pts, onehot)
    finally:
        U.torch, U.knn_point = orig_torch, orig_knn
    loss = torch.nn.CrossEntropyLoss(label_smoothing=0.2)(logits.reshape(-1, cfg["parts"]), labels.reshape(-1))
    loss.backward()
    out = dict(logits=logits.detach().numpy().astype(np.float16) if cfg["N"] > 512 else logits.detach().numpy(),
               loss=np.array([loss.item()]))
    names, norms = [], []
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        if k in FULL_GRADS or k in ("conv3.weight", "conv3.bias", "norm.weight", "label_conv.0.weight",
                                    "propagation.mlp_convs.0.bias", "bn1.weight"):
            out[f"grad::{k}"] = p.grad.numpy()
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms)
    for k, v in ref.state_dict().items():
        if "running_" in k:
            out[f"buf::{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **out)
    print("seg model", name, "loss", out["loss"], "logits", out["logits"].shape, "params with grad", len(names))


def main(what):
    torch.set_num_threads(8)
    if "model" in what:
        for name, cfg in _synth.MODEL_CASES.items():
            run_case(name, cfg)
    if "ft" in what:
        for name, cfg in _synth.FT_CASES.items():
            run_ft_case(name, cfg)
    if "seg" in what:
        for name, cfg in _synth.SEG_CASES.items():
            run_seg_case(name, cfg)
    for w in what:                      # a single part-segmentation case: seg:<name>
        if w.startswith("seg:"):
            run_seg_case(w[4:], _synth.SEG_CASES[w[4:]])


if __name__ == "__main__":
    main(sys.argv[1:] or ["model", "ft", "seg"])
