"""CUDA-event time and achieved HBM bandwidth of the streaming kernels of an encoder layer at the step's shapes
(T = 65536 point-cloud tokens / 36864 image tokens, D = 256, F = 512), against their algorithmic bytes.
cold: a 256 MB buffer is rewritten between runs (L2 flushed);  warm: back to back on the same tensors.

    python tools/bench_elementwise.py [name ...]     names: ln_fwd ln_fwd_add ln_bwd ln_bwd_pos gelu_fwd gelu_bwd dropout_grad
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vipformer_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
PEAK = 6542.7   # MEASURED_PEAKS.json hbm_gbs
which = sys.argv[1:]
reps = int(os.environ.get("REPS", "7"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
seed = torch.tensor([77], device="cuda", dtype=torch.int64)


def timeit(fn, cold):
    ts = []
    for i in range(reps + 2):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


def report(name, T, nbytes, fn):
    c, w = timeit(fn, True), timeit(fn, False)
    print(f"{name:14s} T={T:6d}  {nbytes / 1e6:7.1f} MB   cold {c:7.1f} us {nbytes / c / 1e3:7.0f} GB/s ({nbytes / c / 1e3 / PEAK:4.2f})"
          f"   warm {w:7.1f} us {nbytes / w / 1e3:7.0f} GB/s", flush=True)


for T in (65536, 36864):
    D, Fh, pos_rows = 256, 512, (128 if T == 65536 else 144)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((T, D), device="cuda", generator=g)
    pos = torch.randn((pos_rows, D), device="cuda", generator=g)
    gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    dgam, dbet = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dpos = torch.zeros((pos_rows, D), device="cuda")
    dy = torch.randn((T, D), device="cuda", generator=g).to(BF16)
    dres = torch.randn((T, D), device="cuda", generator=g)
    z = torch.randn((T, Fh), device="cuda", generator=g).to(BF16)
    dh = torch.randn((T, Fh), device="cuda", generator=g).to(BF16)
    cs = torch.zeros(Fh, device="cuda")
    cs2 = torch.zeros(D, device="cuda")
    _, mean, rstd, _ = ops.layernorm_fwd(x, gam, bet)
    cases = {
        "ln_fwd": (T * D * 6, lambda: ops.layernorm_fwd(x, gam, bet)),
        "ln_fwd_add": (T * D * 10, lambda: ops.layernorm_fwd(x, gam, bet, add=pos, want_xsum=True)),
        "ln_bwd": (T * D * 14, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dres, dgamma=dgam, dbeta=dbet)),
        "ln_bwd_nodg": (T * D * 14, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dres)),
        "ln_bwd_pos": (T * D * 14, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dres, dgamma=dgam, dbeta=dbet, dpos=dpos)),
        "gelu_fwd": (T * Fh * 4, lambda: ops.gelu_fwd(z)),
        "gelu_bwd": (T * Fh * 6, lambda: ops.gelu_bwd(dh, z, colsum=cs)),
        "dropout_grad": (T * D * 6, lambda: ops.dropout_grad(dres, 0.5, seed, 7, colsum=cs2)),
    }
    # yardstick: a plain device copy of the same total traffic (torch's copy kernel; what this size can reach at all)
    ca, cb2 = torch.empty(T * D * 7, dtype=torch.uint8, device="cuda"), torch.empty(T * D * 7, dtype=torch.uint8, device="cuda")
    cases["copy(=ln_bwd bytes)"] = (2 * ca.numel(), lambda: cb2.copy_(ca))
    for name, (nb, fn) in cases.items():
        if which and name not in which:
            continue
        report(name, T, nb, fn)

# the BatchNorm column kernels of Group2Emb at the step's size (2.1 M rows x 512 channels, bf16)
if not which or "bn" in which:
    R, C = 512 * 128 * 32, 512
    g = torch.Generator(device="cuda").manual_seed(2)
    xb = torch.randn((R // 8, C), device="cuda", generator=g).to(BF16).repeat(8, 1)
    dyb = torch.randn((R // 8, C), device="cuda", generator=g).to(BF16).repeat(8, 1)
    w, b_ = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    dgw, dgb = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    st = [None]

    def fwd():
        st[0] = ops.bn_forward(xb, w, b_, rm, rv, True, True)[1]

    report("bn_fwd(3 pass)", R, R * C * 2 * 3, fwd)
    report("bn_bwd(2 pass)", R, R * C * 2 * 5, lambda: ops.bn_backward(dyb, xb, st[0], True, dgw, dgb))
    # Group2Emb's own case: 256 channels, followed by the per-patch row sum (separate kernel vs fused)
    x2, dy2 = xb[:, :256].contiguous(), dyb[:, :256].contiguous()
    w2, b2, rm2, rv2 = w[:256].contiguous(), b_[:256].contiguous(), rm[:256].contiguous(), rv[:256].contiguous()
    dg2, db2 = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    st2 = ops.bn_forward(x2, w2, b2, rm2, rv2, True, True)[1]

    def two_kernels():
        d = ops.bn_backward(dy2, x2, st2, True, dg2, db2)
        ops.group_sum(d, R // 32, 32, 256)

    report("bn_bwd+gsum", R, R * 256 * 2 * 6, two_kernels)
    report("bn_bwd_gsum", R, R * 256 * 2 * 5, lambda: ops.bn_backward_gsum(dy2, x2, st2, True, dg2, db2, 32))
