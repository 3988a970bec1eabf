"""Parameter arena: keeps a module tree's fp32 parameters, their fp32 gradients and their bf16 GEMM shadows in three
flat buffers, so that (i) q/k/v (and k/v) projection weights are one contiguous [3D, D] ([2D, D]) operand, (ii) weight
gradients are accumulated in place by the kernels, (iii) gradient all-reduce and AdamW run over ONE buffer
(the DDP + AdamW of pretrain.py:104-105,121-124,209-211), and (iv) bf16 shadows are refreshed by one cast kernel.

Parameters stay ordinary nn.Parameters (same names/shapes as the reference => same state_dict); only their storage is
re-pointed.  Re-pointing is one-time set-up plumbing (torch copy), not hot-path arithmetic.
"""
import torch

from . import ops

_ARENAS = {}   # id(root module) -> Arena
_SHADOW = {}   # id(param) -> (bf16 view, version, data_ptr)


def _align(n, a=64):
    return (n + a - 1) // a * a


class Arena:
    def __init__(self, root):
        self.root = root
        self.gen = 0
        self.flat_p = self.flat_g = self.flat_bf = None
        self.managed = False     # True once an engine owns the update (AdamW rewrites the shadows itself)
        self._versions = None

    def params(self):
        return [p for p in self.root.parameters()]

    # ------------------------------------------------------------------ parameters
    def ensure(self, device):
        # An engine-managed arena was laid out (and verified) when the engine was built and nothing re-points parameter
        # storage afterwards: skip the per-call walk over the module tree (it was > 50 % of the host time of a step).
        if self.managed and self.flat_p is not None and self.flat_p.device == device:
            return
        ps = self.params()
        ok = self.flat_p is not None and self.flat_p.device == device and len(ps) == len(self._offsets)
        if ok:
            base = self.flat_p.data_ptr()
            for p, off in zip(ps, self._offsets):
                if p.data_ptr() != base + 4 * off:
                    ok = False
                    break
        if ok:
            return
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += _align(p.numel())
        flat = torch.zeros(n, dtype=torch.float32, device=device)
        for p, off in zip(ps, offs):
            v = flat[off:off + p.numel()].view(p.shape)
            v.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = v
        self.flat_p, self._offsets, self._n = flat, offs, n
        self.flat_bf = torch.empty(n, dtype=torch.bfloat16, device=device)
        self.flat_g = None
        self._versions = None
        self.gen += 1

    def refresh_shadows(self, force=False):
        if self.managed and not force and self._versions is not None:
            return
        ps = self.params()
        vers = [p._version for p in ps]
        if not force and vers == self._versions:
            return
        ops.cast_bf16(self.flat_p, self.flat_bf)
        self._versions = vers
        for p, off in zip(ps, self._offsets):
            _SHADOW[id(p)] = (self.flat_bf[off:off + p.numel()].view(p.shape), p._version, p.data_ptr())

    # ------------------------------------------------------------------ gradients
    def ensure_grads(self):
        if self.managed and self.flat_g is not None:
            return
        ps = self.params()
        ok = self.flat_g is not None
        if ok:
            base = self.flat_g.data_ptr()
            for p, off in zip(ps, self._offsets):
                if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                    ok = False
                    break
        if ok:
            return
        self.flat_g = torch.empty(self._n, dtype=torch.float32, device=self.flat_p.device)
        ops.zeros_(self.flat_g)
        for p, off in zip(ps, self._offsets):
            old = p.grad
            p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            if old is not None:
                p.grad.copy_(old)
        self.gen += 1

    def zero_grads(self):
        if self.flat_g is not None:
            ops.zeros_(self.flat_g)


def arena_of(root):
    a = _ARENAS.get(id(root))
    if a is None or a.root is not root:
        a = Arena(root)
        _ARENAS[id(root)] = a
    return a


def prepare(root, device):
    """Make `root`'s parameters flat + shadows fresh; returns the arena."""
    a = arena_of(root)
    a.ensure(device)
    a.refresh_shadows()
    return a


def wb(p):
    """bf16 shadow of parameter p (must have been prepared)."""
    s = _SHADOW.get(id(p))
    if s is None or s[2] != p.data_ptr() or s[1] != p._version:
        raise RuntimeError("stale bf16 shadow: call params.prepare(root) before using a parameter in a kernel")
    return s[0]


def adjacent(ts):
    """Tensors laid out back to back (possibly with no padding) -> True."""
    for a, b in zip(ts[:-1], ts[1:]):
        if b.data_ptr() != a.data_ptr() + a.numel() * a.element_size():
            return False
    return True


def cat_view(ts):
    """View of row-concatenated 2-D tensors that are adjacent in memory."""
    if not adjacent(ts):
        raise RuntimeError("parameters are not adjacent in the arena (unexpected module layout)")
    rows = sum(t.shape[0] for t in ts)
    return torch.as_strided(ts[0], (rows, ts[0].shape[1]), (ts[0].shape[1], 1))
