"""Import shim for the REAL reference at /root/reference (build container only).

Used by tests/make_golden.py to generate tests/golden/*.npz, and by the
optional live cross-check tests (skipped when /root/reference is absent, e.g.
on the GPU box).  Three import-time obstacles off the hot path are stubbed:
`imp` (gone in py3.12), fairscale.nn.checkpoint_wrapper, timm DropPath.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "vipformer"))


def load():
    import torch.nn as nn

    if "vipformer" in sys.modules and getattr(sys.modules["vipformer"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["vipformer"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.modules.setdefault("imp", types.ModuleType("imp"))
    fs, fsn = types.ModuleType("fairscale"), types.ModuleType("fairscale.nn")
    fsn.checkpoint_wrapper = lambda m: m
    fs.nn = fsn
    sys.modules.setdefault("fairscale", fs)
    sys.modules.setdefault("fairscale.nn", fsn)

    class _DropPath(nn.Module):
        """timm.models.layers.DropPath (timm is a requirements.txt dependency, absent here) restated from its published
        algorithm: in training, one Bernoulli(1 - p) keep decision per SAMPLE, survivors divided by (1 - p).  The fixture
        generator pins the draw: `scales` [B] (already 0 or 1 / (1 - p)) is assigned per instance before the forward."""

        def __init__(self, p=0.0):
            super().__init__()
            self.p = p
            self.scales = None

        def forward(self, x):
            if self.p == 0.0 or not self.training:
                return x
            if self.scales is None:
                raise RuntimeError("DropPath stub: assign .scales (pinned per-sample draw) before the forward")
            return x * self.scales.to(x.dtype).view(-1, *([1] * (x.ndim - 1)))

    tl = types.ModuleType("timm.models.layers")
    tl.DropPath = _DropPath
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    sys.modules.setdefault("timm.models", types.ModuleType("timm.models"))
    sys.modules.setdefault("timm.models.layers", tl)
    import vipformer  # noqa: F401

    return vipformer
