"""Small host-side runtime shared by the module mirrors: dropout step state, arena lookup, autograd anchors."""
import torch

from . import ops, params

BF16, F32 = torch.bfloat16, torch.float32


class StepState:
    """Device-resident per-step state shared by all dropout sites: state[0] = optimizer step, state[1] = seed.
    Kernels read the seed through a device pointer, so a captured CUDA graph sees a fresh seed every replay."""
    _by_device = {}

    @classmethod
    def get(cls, device):
        device = torch.device(device)
        st = cls._by_device.get(device)
        if st is None:
            st = torch.tensor([0, 0x1E3779B97F4A7C15], dtype=torch.int64, device=device)
            cls._by_device[device] = st
        return st

    @classmethod
    def seed_ptr(cls, device):
        return cls.get(device)[1:]

    @classmethod
    def step_ptr(cls, device):
        return cls.get(device)[:1]

    @classmethod
    def advance(cls, device):
        ops.step_advance(cls.get(device))


def _dev(device):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def manual_seed(seed, device=None):
    """Seed the counter-based dropout stream (the backward pass regenerates masks from the same seed)."""
    StepState.get(_dev(device))[1] = int(seed) & 0x7FFFFFFFFFFFFFFF


def advance_dropout_seed(device=None):
    """Call once per training step (the engine does) so every step draws fresh dropout masks."""
    StepState.advance(_dev(device))


def root_prepare(module, device):
    """Flatten + shadow the parameters of the outermost vipformer_b200 module this call came through."""
    root = module.__dict__.get("_vpf_root") or module
    return params.prepare(root, device)


def set_root(root):
    for m in root.modules():
        if m is not root and m.__dict__.get("_vpf_root") is not root:
            object.__setattr__(m, "_vpf_root", root)


def anchor(module):
    """A parameter passed through autograd so that backward runs even when no data input requires grad."""
    for p in module.parameters():
        if p.requires_grad:
            return p
    return None


def as_f32_2d(t, cols):
    t = t.reshape(-1, cols)
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


def fragment_error(name):
    raise RuntimeError(f"{name} is a fragment of a fused layer kernel sequence in vipformer_b200; call the enclosing "
                       "CrossAttentionLayer / SelfAttentionLayer / Encoder instead")


def bump(bn):
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
