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


# ------------------------------------------------------------------------------------------ model fixtures
MODEL_CASES = {
    # name: dict(D, H, n_sa, G, S, N, MR, b (pairs), img)
    "small": dict(D=128, H=2, n_sa=2, G=32, S=8, N=128, MR=2, b=6, img=144, patch=12, seed=21),
    "cfgA": dict(D=256, H=4, n_sa=8, G=128, S=32, N=2048, MR=2, b=4, img=144, patch=12, seed=22),
}


def build_models(cfg, atten_drop=0.0, mlp_drop=0.0, pkg="vipformer_b200"):
    """Construct (pc_model, img_model) with the kwargs of the reference's utils.build_model (utils.py:115-148),
    from `pkg` = 'vipformer_b200' (the product mirror) or 'vipformer' (the real reference, build container only)."""
    import importlib

    import torch

    pcmod = importlib.import_module(pkg + ".model.pointcloud")
    torch.manual_seed(cfg["seed"])
    ad = pcmod.PointCloudInputAdapter(pointcloud_shape=(cfg["N"], 3), num_input_channels=cfg["D"])
    common = dict(num_latent_channels=cfg["D"], num_cross_attention_layers=1, num_cross_attention_heads=cfg["H"],
                  num_self_attention_layers=cfg["n_sa"], num_self_attention_heads=cfg["H"],
                  mlp_widen_factor=cfg["MR"], max_dpr=0.0, atten_drop=atten_drop, mlp_drop=mlp_drop, modal_prior=True)
    pc = pcmod.CrossFormer_pc_mp(input_adapter=ad, num_latents=cfg["G"], group_size=cfg["S"], **common)
    im = pcmod.CrossFormer_img_mp(img_height=cfg["img"], img_width=cfg["img"], patch_size=cfg["patch"], **common)
    return pc, im


FT_CASES = {
    # fine-tune classification fixtures (BASELINE configs[3]: ScanObjectNN-shaped, 15 classes)
    "ft_small": dict(D=128, H=2, n_sa=2, G=32, S=8, N=128, MR=2, b=10, classes=15, seed=61),
    "ft_cfgA": dict(D=256, H=4, n_sa=8, G=128, S=32, N=2048, MR=2, b=8, classes=15, seed=62),
}


def build_ft_model(cfg, atten_drop=0.0, mlp_drop=0.0, pkg="vipformer_b200"):
    """CrossFormer_pc_mp_ft with the kwargs of the reference's utils.build_ft_cls (utils.py:203-224)."""
    import importlib

    import torch

    pcmod = importlib.import_module(pkg + ".model.pointcloud")
    part = importlib.import_module(pkg + ".model.pointcloud.partseg")
    torch.manual_seed(cfg["seed"])
    ad = pcmod.PointCloudInputAdapter(pointcloud_shape=(cfg["N"], 3), num_input_channels=cfg["D"])
    return part.CrossFormer_pc_mp_ft(input_adapter=ad, num_latents=cfg["G"], num_latent_channels=cfg["D"], group_size=cfg["S"],
                                     num_cross_attention_layers=1, num_cross_attention_heads=cfg["H"],
                                     num_self_attention_layers=cfg["n_sa"], num_self_attention_heads=cfg["H"],
                                     mlp_widen_factor=cfg["MR"], max_dpr=0.0, atten_drop=atten_drop, mlp_drop=mlp_drop,
                                     modal_prior=True, num_obj_classes=cfg["classes"])


def ft_inputs(cfg):
    import torch

    pts = make_clouds("randn", cfg["b"], cfg["N"], cfg["seed"] + 1)
    start = make_start(cfg["b"], cfg["N"], cfg["seed"])
    g = torch.Generator().manual_seed(cfg["seed"] + 3)
    labels = torch.randint(0, cfg["classes"], (cfg["b"],), generator=g)
    return torch.from_numpy(pts), start, labels


SEG_CASES = {
    # part-segmentation fixtures (ft_partseg.py: 16 object classes, 50 part classes)
    "seg_small": dict(D=128, H=2, n_sa=4, G=32, S=8, N=128, MR=2, b=4, layer_idx=[1, 2, 4], parts=50, seed=71),
    "seg_cfgA": dict(D=256, H=4, n_sa=8, G=128, S=32, N=2048, MR=2, b=4, layer_idx=[2, 5, 8], parts=50, seed=72),
    # DropPath on (the reference's own default for this model is max_dpr = 0.1): 8 samples so that several are dropped
    "seg_small_dpr": dict(D=128, H=2, n_sa=4, G=32, S=8, N=128, MR=2, b=8, layer_idx=[1, 2, 4], parts=50, seed=73, max_dpr=0.45),
}
DPR_SEED = 0x5EED0D9A          # seed of the pinned DropPath draws of the fixtures (oracle/rng.py droppath_scales)


def seg_op_bases(cfg):
    """Dropout / DropPath site ids the fixtures and the oracle use for the part-segmentation encoder: layer i -> 8 (i + 2)."""
    ob = {"seg.encoder.cross_attn_1": 8}
    for i in range(cfg["n_sa"]):
        ob[f"seg.encoder.sa_layers.{i}"] = 8 * (i + 2)
    return ob


def seg_drop_path(cfg):
    """{layer key: DropPath rate}: torch.linspace(0, max_dpr, n_sa) as partseg.py:372 does."""
    import torch

    rates = [x.item() for x in torch.linspace(0, cfg.get("max_dpr", 0.0), cfg["n_sa"])]
    return {f"seg.encoder.sa_layers.{i}": r for i, r in enumerate(rates) if r > 0.0}


def build_seg_model(cfg, atten_drop=0.0, mlp_drop=0.0, pkg="vipformer_b200"):
    """CrossFormer_partseg with the kwargs of the reference's utils.build_ft_partseg (utils.py:277-298); max_dpr = 0."""
    import importlib

    import torch

    pcmod = importlib.import_module(pkg + ".model.pointcloud")
    part = importlib.import_module(pkg + ".model.pointcloud.partseg")
    torch.manual_seed(cfg["seed"])
    ad = pcmod.PointCloudInputAdapter(pointcloud_shape=(cfg["N"], 3), num_input_channels=cfg["D"])
    return part.CrossFormer_partseg(input_adapter=ad, num_latents=cfg["G"], num_latent_channels=cfg["D"], group_size=cfg["S"],
                                    num_cross_attention_layers=1, num_cross_attention_heads=cfg["H"],
                                    num_self_attention_layers=cfg["n_sa"], num_self_attention_heads=cfg["H"],
                                    mlp_widen_factor=cfg["MR"], max_dpr=cfg.get("max_dpr", 0.0), atten_drop=atten_drop,
                                    mlp_drop=mlp_drop, layer_idx=list(cfg["layer_idx"]), num_part_classes=cfg["parts"])


def seg_inputs(cfg):
    """-> (pts [b,N,3], FPS start [b], one-hot object class [b,16], part labels [b,N])  (ft_partseg.py:146-160)."""
    import torch

    pts = make_clouds("randn", cfg["b"], cfg["N"], cfg["seed"] + 1)
    start = make_start(cfg["b"], cfg["N"], cfg["seed"])
    g = torch.Generator().manual_seed(cfg["seed"] + 3)
    obj = torch.randint(0, 16, (cfg["b"],), generator=g)
    onehot = torch.nn.functional.one_hot(obj, 16).float()
    labels = torch.randint(0, cfg["parts"], (cfg["b"], cfg["N"]), generator=g)
    return torch.from_numpy(pts), start, onehot, labels


def perturb_state_dict(sd, seed):
    """Deterministic perturbation so LayerNorm/BatchNorm affine terms and biases are not at their trivial init."""
    import torch

    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.dtype.is_floating_point:
            noise = torch.randn(v.shape, generator=g)
            if k.endswith("running_var"):
                v = v + 0.1 * noise.abs()
            else:
                v = v + 0.02 * noise
        out[k] = v.clone()
    for k in list(out.keys()):   # cross_attn_1 / cross_attn_n alias ONE module (partseg.py:295-300): keep them equal
        if "cross_attn_n." in k:
            out[k] = out[k.replace("cross_attn_n.", "cross_attn_1.")].clone()
    return out


def model_inputs(cfg):
    import torch

    b = cfg["b"]
    pts = np.concatenate([make_clouds("randn", b, cfg["N"], cfg["seed"] + 1), make_clouds("randn", b, cfg["N"], cfg["seed"] + 2)], 0)
    start = make_start(2 * b, cfg["N"], cfg["seed"])
    g = torch.Generator().manual_seed(cfg["seed"] + 3)
    imgs = torch.randn((b, cfg["img"], cfg["img"], 3), generator=g)
    return torch.from_numpy(pts), start, imgs
