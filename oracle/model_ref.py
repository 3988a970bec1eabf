"""Plain PyTorch fp32/fp64 CPU restatement of the floating-point half of the hot path -- TEST INFRASTRUCTURE ONLY.

Functional style over a reference-layout state_dict (same keys as the reference's modules), so the
same weights drive this oracle, the real reference (tests/make_golden_model.py, build container only)
and the CUDA product.  Autograd of this restatement is the gradient oracle.

Each function cites the reference lines it restates (paths relative to the upstream repo root).
Pinned against the real reference by tests/golden/model_*.npz (tests/test_oracle_golden.py).
`ntxent` restates the un-vendored lightly==1.1.21 package => that function is "parity unpinned"
(no reference source or test vector exists in-tree).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import tokenizer as T


# ---- discrete-choice pinning (test infrastructure) ---------------------------------------------------------------
# The hot path contains DISCRETE choices -- ReLU masks (utils.py:156,163; classifier.py:34; partseg.py:521,524) and
# max-pool arg-maxes (utils.py:180,188; partseg.py:547) -- which the bf16 rounding of the product's forward
# activations can flip.  A flip moves a whole gradient row, so an unpinned gradient comparison measures the flip rate,
# not the kernels.  `with choices(pins) as ch:` makes this oracle take the masks / arg-max indices it is handed (the
# ones the product's forward actually used) instead of its own, and records its own in `ch.rec` so the flip rate can be
# reported separately.  With pins = None the oracle is unchanged (and is what the golden files pin).
class _Choices:
    def __init__(self, pins):
        self.pins, self.rec, self.prefix = (pins or {}), {}, ""


_CH = None


class choices:
    def __init__(self, pins=None):
        self.ch = _Choices(pins)

    def __enter__(self):
        global _CH
        self.prev, _CH = _CH, self.ch
        return self.ch

    def __exit__(self, *a):
        global _CH
        _CH = self.prev


def _relu(x, name):
    if _CH is None:
        return F.relu(x)
    name = _CH.prefix + name
    _CH.rec[name] = (x > 0).detach()
    if name in _CH.pins:
        return x * _CH.pins[name].reshape(x.shape).to(x.dtype)
    return F.relu(x)


def _max1(x, name):
    """max over dim 1 of x [A, S, C] -> [A, C] (torch.max(dim): values only are used by the reference)."""
    val, idx = x.max(1)
    if _CH is None:
        return val
    name = _CH.prefix + name
    _CH.rec[name] = idx.detach()
    if name in _CH.pins:
        return x.gather(1, _CH.pins[name].reshape(idx.shape).long().unsqueeze(1)).squeeze(1)
    return val


def _leaky(x, slope, name):
    if _CH is None:
        return F.leaky_relu(x, slope)
    name = _CH.prefix + name
    _CH.rec[name] = (x > 0).detach()
    if name in _CH.pins:
        m = _CH.pins[name].reshape(x.shape).to(x.dtype)
        return x * (m + slope * (1 - m))
    return F.leaky_relu(x, slope)


_PREFIX = [""]


class _prefix:
    def __init__(self, p):
        self.p = p

    def __enter__(self):
        self.prev = _PREFIX[0]
        _PREFIX[0] = self.p
        if _CH is not None:
            _CH.prefix = self.p

    def __exit__(self, *a):
        _PREFIX[0] = self.prev
        if _CH is not None:
            _CH.prefix = self.prev


# ---- dropout-mask injection (test infrastructure) ------------------------------------------------------------------
# The reference draws nn.Dropout masks from torch's RNG (partseg.py:81,208-213); the product's masks are counter-based
# (oracle/rng.py restates them).  `with dropout(seed, op_bases, atten_drop, mlp_drop):` makes this oracle apply exactly
# the product's keep-masks: attention-probability dropout after the softmax (partseg.py:81) and Residual dropout on
# f(x) before the skip connection (partseg.py:208-213; rates per partseg.py:165-166,186-187).
_DROP = None


class dropout:
    def __init__(self, seed, op_bases, atten_drop, mlp_drop, drop_path=None):
        """op_bases: {"pc.encoder.cross_attn_1": id, "pc.encoder.sa_layers.0": id, ..., "img...."} (the product's).
        drop_path: {layer key (as in op_bases): DropPath rate} -- per-sample scales of oracle/rng.py droppath_scales."""
        self.cfg = dict(seed=int(seed), op_bases=op_bases, atten_drop=float(atten_drop), mlp_drop=float(mlp_drop),
                        drop_path=dict(drop_path or {}))

    def __enter__(self):
        global _DROP
        self.prev, _DROP = _DROP, self.cfg
        return self

    def __exit__(self, *a):
        global _DROP
        _DROP = self.prev


def _drop_attn(attn, layer_key, p):
    if _DROP is None or p <= 0.0:
        return attn
    from . import rng as R
    B, H, Lq, Lk = attn.shape
    m = R.attention_keep(_DROP["seed"], _DROP["op_bases"][_PREFIX[0] + layer_key], p, B * H, Lq, Lk)
    return attn * torch.from_numpy(m).view(B, H, Lq, Lk).to(attn.dtype)


def _drop_resid(y, layer_key, which, p):
    """y [B, L, D]: output of f(x) inside Residual; which = 1 (attention residual) or 2 (MLP residual)."""
    if _DROP is None or p <= 0.0:
        return y
    from . import rng as R
    B, L, D = y.shape
    m = R.residual_keep(_DROP["seed"], _DROP["op_bases"][_PREFIX[0] + layer_key] + which, p, B * L, D)
    return y * torch.from_numpy(m).view(B, L, D).to(y.dtype)


def _drop_path(x, layer_key, which):
    """DropPath of Residual number `which` (1 attention, 2 MLP) of a layer, partseg.py:206,212: applied to the WHOLE sum
    dropout(f(x)) + x (the reference's quirk), one keep decision per sample, survivors / (1 - p)."""
    if _DROP is None:
        return x
    p = _DROP["drop_path"].get(_PREFIX[0] + layer_key, 0.0)
    if p <= 0.0:
        return x
    from . import rng as R
    s = R.droppath_scales(_DROP["seed"], _DROP["op_bases"][_PREFIX[0] + layer_key] + 2 + which, p, x.shape[0])
    return x * torch.from_numpy(s).view(-1, 1, 1).to(x.dtype)


def _lin(sd, k, x, bias=True):
    return F.linear(x, sd[k + ".weight"], sd[k + ".bias"] if bias and (k + ".bias") in sd else None)


def _ln(sd, k, x):
    return F.layer_norm(x, (x.shape[-1],), sd[k + ".weight"], sd[k + ".bias"], 1e-5)


def _bn(sd, k, x, training, running_out=None):
    """nn.BatchNorm1d on [R, C] (or [B, C, L] flattened by the caller): train = batch stats (biased var), eps 1e-5."""
    if training:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        if running_out is not None:   # momentum 0.1, unbiased variance into running_var
            running_out[k + ".running_mean"] = 0.9 * sd[k + ".running_mean"] + 0.1 * mean.detach()
            running_out[k + ".running_var"] = 0.9 * sd[k + ".running_var"] + 0.1 * x.var(0, unbiased=True).detach()
    else:
        mean, var = sd[k + ".running_mean"], sd[k + ".running_var"]
    return (x - mean) / torch.sqrt(var + 1e-5) * sd[k + ".weight"] + sd[k + ".bias"]


def mha(sd, k, xq, xkv, H, layer_key=None, p_attn=0.0):
    """MultiHeadAttention.forward, partseg.py:53-86 (pad_mask None; dropout only through `with dropout(...)`)."""
    q, kk, v = _lin(sd, k + ".q_proj", xq, False), _lin(sd, k + ".k_proj", xkv, False), _lin(sd, k + ".v_proj", xkv, False)
    B, Lq, D = q.shape
    Lk = kk.shape[1]
    dh = D // H
    q = q.view(B, Lq, H, dh).transpose(1, 2)
    kk = kk.view(B, Lk, H, dh).transpose(1, 2)
    v = v.view(B, Lk, H, dh).transpose(1, 2)
    attn = (q @ kk.transpose(-1, -2)) * dh ** -0.5
    attn = _drop_attn(attn.softmax(-1), layer_key, p_attn)
    o = (attn @ v).transpose(1, 2).reshape(B, Lq, D)
    return _lin(sd, k + ".o_proj", o)


def mlp(sd, k, x):
    """MLP, partseg.py:191-198: LN, Linear, exact GELU, Linear."""
    return _lin(sd, k + ".3", F.gelu(_lin(sd, k + ".1", _ln(sd, k + ".0", x))))


def ca_layer(sd, k, xq, xkv, H):
    """CrossAttentionLayer, partseg.py:144-167 with Residual (:201-213); dropout rates of :165-166."""
    a = k + ".0.module"
    pa, pm = (_DROP["atten_drop"], _DROP["mlp_drop"]) if _DROP else (0.0, 0.0)
    y = mha(sd, a + ".attention", _ln(sd, a + ".q_norm", xq), _ln(sd, a + ".kv_norm", xkv), H, k, pa)
    x = _drop_resid(y, k, 1, pa) + xq
    return _drop_resid(mlp(sd, k + ".1.module", x), k, 2, pm) + x


def sa_layer(sd, k, x, H):
    """SelfAttentionLayer, partseg.py:170-188 (both residuals use mlp_drop, :186-187)."""
    a = k + ".0.module"
    pa, pm = (_DROP["atten_drop"], _DROP["mlp_drop"]) if _DROP else (0.0, 0.0)
    xn = _ln(sd, a + ".norm", x)
    x = _drop_path(_drop_resid(mha(sd, a + ".attention", xn, xn, H, k, pa), k, 1, pm) + x, k, 1)
    return _drop_path(_drop_resid(mlp(sd, k + ".1.module", x), k, 2, pm) + x, k, 2)


def encoder(sd, k, group_embs, pos_embs, pts_embs, H, n_sa):
    """Encoder.forward, partseg.py:314-342 (one cross-attention layer, modal_prior=True)."""
    x = ca_layer(sd, k + ".cross_attn_1", group_embs + pos_embs, pts_embs, H)
    for i in range(n_sa):
        x = sa_layer(sd, f"{k}.sa_layers.{i}", x + pos_embs, H)
    return x


def group2emb(sd, k, nb, training, running_out=None):
    """Group2Emb.forward, utils.py:168-189.  nb [B,G,S,3] -> [B,G,D]."""
    bs, g, n, _ = nb.shape
    x = nb.reshape(bs * g * n, 3)
    x = F.linear(x, sd[k + ".first_conv.0.weight"][:, :, 0], sd[k + ".first_conv.0.bias"])
    x = _relu(_bn(sd, k + ".first_conv.1", x, training, running_out), "g2e.relu1")
    x = F.linear(x, sd[k + ".first_conv.3.weight"][:, :, 0], sd[k + ".first_conv.3.bias"])  # [R,128]
    xg = _max1(x.view(bs * g, n, 128), "g2e.max2").unsqueeze(1).expand(-1, n, -1)
    x = torch.cat([xg, x.view(bs * g, n, 128)], -1).reshape(bs * g * n, 256)
    x = F.linear(x, sd[k + ".second_conv.0.weight"][:, :, 0], sd[k + ".second_conv.0.bias"])
    x = _relu(_bn(sd, k + ".second_conv.1", x, training, running_out), "g2e.relu3")
    x = F.linear(x, sd[k + ".second_conv.3.weight"][:, :, 0], sd[k + ".second_conv.3.bias"])
    D = x.shape[-1]
    return _max1(x.view(bs * g, n, D), "g2e.max4").view(bs, g, D)


def input_adapter(sd, k, pts):
    """PointCloudInputAdapter.forward, classifier.py:31-50."""
    m = k + ".point_mlp"
    return _lin(sd, m + ".3", _relu(_ln(sd, m + ".1", _lin(sd, m + ".0", pts)), "adapter.relu"))


def position_emb(sd, k, center):
    """partseg.py:498-501."""
    return _lin(sd, k + ".2", F.gelu(_lin(sd, k + ".0", center)))


def latent_head(sd, k, x, training, running_out=None):
    """partseg.py:519-525 on backbone feats [B,2D]."""
    x = _relu(_bn(sd, k + ".0", x, training, running_out), "head.relu1")
    x = F.linear(x, sd[k + ".2.weight"])
    x = _relu(_bn(sd, k + ".3", x, training, running_out), "head.relu2")
    return F.linear(x, sd[k + ".5.weight"])


def pc_forward(sd, pts, start_idx, G, S, H, n_sa, training=True, running_out=None, tokenizer=None):
    """CrossFormer_pc_mp.forward, partseg.py:527-550, with the tokenizer pinned as in oracle/tokenizer_oracle.c
    (`tokenizer`: alternative callable (pts, G, S, start_idx) -> (neighbors, centers), e.g. oracle.tokenizer_torch on a GPU)."""
    with _prefix("pc."):
        pts_embs = input_adapter(sd, "input_adapter", pts)
        if tokenizer is not None:
            nb, ce = tokenizer(pts.detach(), G, S, start_idx)
        else:
            nb, ce = T.divide_patches(pts.detach().numpy(), G, S, np.asarray(start_idx))
            nb, ce = torch.from_numpy(nb).to(pts.dtype), torch.from_numpy(ce).to(pts.dtype)
        group_embs = group2emb(sd, "group2emb", nb, training, running_out)
        pos_embs = position_emb(sd, "position_emb", ce)
        x = encoder(sd, "encoder", group_embs, pos_embs, pts_embs, H, n_sa)
        backbone = torch.cat([_max1(x, "pool.max"), x.mean(1)], 1)
        return latent_head(sd, "latent_head", backbone, training, running_out), backbone


def finetune_head(sd, k, x, training, running_out=None):
    """partseg.py:573-582 on backbone feats [B,2D]: 3 x {BatchNorm1d, ReLU, Linear (with bias)}."""
    for i, (bn, fc) in enumerate(((0, 2), (3, 5), (6, 8))):
        x = _relu(_bn(sd, f"{k}.{bn}", x, training, running_out), f"cls.relu{i + 1}")
        x = F.linear(x, sd[f"{k}.{fc}.weight"], sd[f"{k}.{fc}.bias"])
    return x


def pc_ft_forward(sd, pts, start_idx, G, S, H, n_sa, training=True, running_out=None, tokenizer=None):
    """CrossFormer_pc_mp_ft.forward, partseg.py:584-605 -> logits [B, num_obj_classes]."""
    with _prefix("pc."):
        pts_embs = input_adapter(sd, "input_adapter", pts)
        if tokenizer is not None:
            nb, ce = tokenizer(pts.detach(), G, S, start_idx)
        else:
            nb, ce = T.divide_patches(pts.detach().numpy(), G, S, np.asarray(start_idx))
            nb, ce = torch.from_numpy(nb).to(pts.dtype), torch.from_numpy(ce).to(pts.dtype)
        group_embs = group2emb(sd, "group2emb", nb, training, running_out)
        pos_embs = position_emb(sd, "position_emb", ce)
        x = encoder(sd, "encoder", group_embs, pos_embs, pts_embs, H, n_sa)
        backbone = torch.cat([_max1(x, "pool.max"), x.mean(1)], 1)
        return finetune_head(sd, "finetune_head", backbone, training, running_out)


def encoder_taps(sd, k, group_embs, pos_embs, pts_embs, H, n_sa, layer_idx):
    """Encoder.forward with modal_prior=False, partseg.py:314-341: the outputs of self-attention layers i+1 in layer_idx."""
    x = ca_layer(sd, k + ".cross_attn_1", group_embs + pos_embs, pts_embs, H)
    feats = []
    for i in range(n_sa):
        x = sa_layer(sd, f"{k}.sa_layers.{i}", x + pos_embs, H)
        if i + 1 in layer_idx:
            feats.append(x)
    return feats


def three_nn(xyz1, xyz2):
    """PointNetFeaturePropagation's neighbour search, utils.py:223-229 with square_distance (utils.py:138-140) written out:
    xyz1 [B,N,3], xyz2 [B,S,3] -> (idx [B,N,3] int64, weight [B,N,3]).  Pin name "nn3" replaces idx (weights follow it)."""
    B, N, _ = xyz1.shape
    dist = -2 * torch.matmul(xyz1, xyz2.permute(0, 2, 1))
    dist = dist + torch.sum(xyz1 ** 2, -1).view(B, N, 1)
    dist = dist + torch.sum(xyz2 ** 2, -1).view(B, 1, -1)
    d, idx = dist.sort(dim=-1)
    d, idx = d[:, :, :3], idx[:, :, :3]
    if _CH is not None:
        name = _CH.prefix + "nn3"
        _CH.rec[name] = idx.detach()
        if name in _CH.pins:
            idx = _CH.pins[name].reshape(idx.shape).long()
            d = dist.gather(2, idx)
    r = 1.0 / (d + 1e-8)
    return idx, r / r.sum(2, keepdim=True)


def feature_propagation(sd, k, pts, center, x, training, running_out=None):
    """PointNetFeaturePropagation.forward, utils.py:205-242, rows x channels: pts [B,N,3], center [B,S,3], x [B,S,C]
    (points1 = the xyz of every point, partseg.py:450) -> [B*N, 1024]."""
    B, N, _ = pts.shape
    idx, w = three_nn(pts, center)
    gathered = torch.stack([x[b][idx[b]] for b in range(B)])                  # index_points: [B, N, 3, C]
    interp = (gathered * w.view(B, N, 3, 1)).sum(2)
    f = torch.cat([pts, interp], -1).reshape(B * N, -1)                       # cat([points1, interpolated]), utils.py:234
    for i in range(2):
        f = F.linear(f, sd[f"{k}.mlp_convs.{i}.weight"][:, :, 0], sd[f"{k}.mlp_convs.{i}.bias"])
        f = _relu(_bn(sd, f"{k}.mlp_bns.{i}", f, training, running_out), f"seg.relu_p{i + 1}")
    return f


def partseg_forward(sd, pts, cls_onehot, start_idx, G, S, H, n_sa, layer_idx, training=True, running_out=None,
                    tokenizer=None):
    """CrossFormer_partseg.forward, partseg.py:407-470 (max_dpr = 0: DropPath is the identity) -> [B, N, num_part_classes].
    Conv1d(kernel 1) layers are written as linear maps over rows = points; dp1's mask comes from `with dropout(...)` under the
    key "seg.dp1" (rows x 512 channels, element index row * 512 + channel), else it is the identity."""
    with _prefix("seg."):
        B, N, _ = pts.shape
        pts_embs = input_adapter(sd, "input_adapter", pts)
        if tokenizer is not None:
            nb, ce = tokenizer(pts.detach(), G, S, start_idx)
        else:
            nb, ce = T.divide_patches(pts.detach().numpy(), G, S, np.asarray(start_idx))
            nb, ce = torch.from_numpy(nb).to(pts.dtype), torch.from_numpy(ce).to(pts.dtype)
        group_embs = group2emb(sd, "group2emb", nb, training, running_out)
        pos_embs = position_emb(sd, "position_emb", ce)
        feats = encoder_taps(sd, "encoder", group_embs, pos_embs, pts_embs, H, n_sa, layer_idx)
        x = torch.cat([_ln(sd, "norm", f) for f in feats], -1)                # [B, G, k*D]  (:422-428)
        x_max, x_avg = _max1(x, "seg.max"), x.mean(1)                         # :430-434
        lab = F.linear(cls_onehot.view(B, 16).to(pts.dtype), sd["label_conv.0.weight"][:, :, 0])
        lab = _leaky(_bn(sd, "label_conv.1", lab, training, running_out), 0.2, "seg.leaky")      # :391-393, :441-443
        glob = torch.cat([x_max, x_avg, lab], 1)                              # [B, 2kD + 64], repeated over the points
        f0 = feature_propagation(sd, "propagation", pts, ce, x, training, running_out)           # [B*N, 1024]
        h = torch.cat([f0, glob.unsqueeze(1).expand(-1, N, -1).reshape(B * N, -1)], 1)           # :452
        h = F.linear(h, sd["conv1.weight"][:, :, 0], sd["conv1.bias"])
        h = _relu(_bn(sd, "bn1", h, training, running_out), "seg.relu1")
        if _DROP is not None and training and "seg.dp1" in _DROP["op_bases"]:
            from . import rng as R
            m = R.residual_keep(_DROP["seed"], _DROP["op_bases"]["seg.dp1"], 0.5, B * N, 512)
            h = h * torch.from_numpy(m).to(h.dtype)
        h = F.linear(h, sd["conv2.weight"][:, :, 0], sd["conv2.bias"])
        h = _relu(_bn(sd, "bn2", h, training, running_out), "seg.relu2")
        h = F.linear(h, sd["conv3.weight"][:, :, 0], sd["conv3.bias"])
        return h.view(B, N, -1)


def cross_entropy_ls(logits, labels, eps=0.2):
    """torch.nn.CrossEntropyLoss(label_smoothing=eps), ft_cls.py:145 -- written out (mean reduction)."""
    logp = F.log_softmax(logits, dim=1)
    nll = -logp.gather(1, labels.view(-1, 1)).squeeze(1)
    smooth = -logp.mean(1)
    return ((1 - eps) * nll + eps * smooth).mean()


def patch2emb(sd, k, imgs, patch):
    """partseg.py:631-634: Rearrange 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)' + Linear."""
    B, Hh, Ww, C = imgs.shape
    x = imgs.view(B, Hh // patch, patch, Ww // patch, patch, C).permute(0, 1, 3, 2, 4, 5).reshape(B, -1, patch * patch * C)
    return _lin(sd, k + ".1", x)


def img_forward(sd, imgs, patch, H, n_sa, training=True, running_out=None):
    """CrossFormer_img_mp.forward, partseg.py:661-680."""
    with _prefix("img."):
        e = patch2emb(sd, "patch2emb", imgs, patch)
        x = encoder(sd, "encoder", e, sd["position_emb"], e, H, n_sa)
        backbone = torch.cat([_max1(x, "pool.max"), x.mean(1)], 1)
        return latent_head(sd, "latent_head", backbone, training, running_out), backbone


def ntxent(out0, out1, temperature=0.1):
    """lightly==1.1.21 lightly/loss/ntx_ent_loss.py NTXentLoss.forward (memory bank off) -- PARITY UNPINNED.
    Call sites: pretrain.py:155,196,202."""
    out0, out1 = F.normalize(out0, dim=1), F.normalize(out1, dim=1)
    b = out0.shape[0]
    z = torch.cat([out0, out1], 0)
    logits = torch.einsum("nc,mc->nm", z, z) / temperature
    logits = logits[~torch.eye(2 * b, dtype=torch.bool, device=z.device)].view(2 * b, -1)
    labels = torch.cat([torch.arange(b, device=z.device) + b - 1, torch.arange(b, device=z.device)])
    return F.cross_entropy(logits, labels)


def ntxent_closed_form(out0, out1, temperature=0.1):
    """mean_i [ logsumexp_{j != i} s_ij - s_{i,pos(i)} ] -- the algebraic definition the restatement must equal."""
    z = torch.cat([F.normalize(out0, dim=1), F.normalize(out1, dim=1)], 0).double()
    n = z.shape[0]
    s = z @ z.t() / temperature
    s_masked = s.masked_fill(torch.eye(n, dtype=torch.bool), -float("inf"))
    pos = torch.cat([torch.arange(n // 2) + n // 2, torch.arange(n // 2)])
    return (torch.logsumexp(s_masked, 1) - s[torch.arange(n), pos]).mean()


def pretrain_loss(pc_feats, img_feats, cmid_weight=1.0, temperature=0.1):
    """Loss composition of pretrain.py:189-207 (modality 'both')."""
    b = pc_feats.shape[0] // 2
    t1, t2 = pc_feats[:b], pc_feats[b:]
    imid = ntxent(t1, t2, temperature)
    cmid = ntxent((t1 + t2) / 2, img_feats, temperature)
    return imid + cmid_weight * cmid, imid, cmid
