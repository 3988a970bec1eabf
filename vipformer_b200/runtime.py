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


# ---- per-forward dropout epoch (drop-in path, no engine) ---------------------------------------------------------
# Dropout masks are keep(e) = hash(seed, op_id, element).  PretrainEngine advances the device-resident seed once per
# step.  On the plain drop-in path (model(x); loss.backward(); optimizer.step()) nobody does, so every training-mode
# encoder forward folds a host-side epoch counter into the op ids instead: consecutive forwards draw different masks
# (nn.Dropout semantics, partseg.py:81,208-213) and the backward pass, which reads the op ids saved at forward time,
# regenerates exactly the forward's masks.
_EPOCH = [0]
EPOCH_STRIDE = 1 << 12


def next_op_offset(managed):
    """Offset added to a layer's dropout op ids for this forward call (0 under an engine-managed arena, whose seed
    advances on the device so that a captured CUDA graph stays valid)."""
    if managed:
        return 0
    _EPOCH[0] = (_EPOCH[0] + 1) & 0xFFFFF
    return _EPOCH[0] * EPOCH_STRIDE


# ---- auxiliary stream of a device (tokenizer next to the input adapter, CrossFormer_pc_mp._tokens)
_AUX = {}


def aux_stream(device):
    import torch

    key = torch.device(device).index
    if key not in _AUX:
        _AUX[key] = torch.cuda.Stream(device=device)
    return _AUX[key]


# ---- test tap: when set to a dict, block forwards drop their saved context here (discrete choices for parity tests)
TAP = None


def tap(name, value):
    if TAP is not None:
        TAP.setdefault(name, []).append(value)


def _dev(device):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def manual_seed(seed, device=None):
    """Seed the counter-based dropout stream (the backward pass regenerates masks from the same seed)."""
    StepState.get(_dev(device))[1] = int(seed) & 0x7FFFFFFFFFFFFFFF
    _EPOCH[0] = 0


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
