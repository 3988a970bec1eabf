"""Marginal in-graph cost of each kernel family of the pre-training step: rebuild the engine with one family of C-ABI
entries turned into no-ops (results are garbage, timing is not: no kernel on this path branches on data except the
attention's lazy rescale), capture the step graph, time it, report baseline - ablated.  ncu's per-launch times are
cold-cache and serialised; this is what a kernel costs inside the captured two-stream step.
CAVEAT: the skipped kernels leave NaN garbage behind, which changes later kernels' data-dependent
branches -- the numbers are hints for where to look, not a cost model; `--only baseline` (nothing skipped) is the reliable
use: same-box A/B timing of a build under the VPF_* environment switches.

    python tools/ablate_step.py [--pairs 256] [--steps 10] [--config A]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FAMILIES = {
    "baseline": [],
    "gelu_fwd": ["vpf_gelu_fwd"],
    "gelu_bwd": ["vpf_gelu_bwd"],
    "dropout_grad": ["vpf_dropout_grad"],
    "layernorm_fwd": ["vpf_layernorm_fwd"],
    "layernorm_bwd": ["vpf_layernorm_bwd"],
    "attention_fwd": ["vpf_attention_fwd"],
    "attention_bwd": ["vpf_attention_bwd"],
    "gemm_wgrad(atomic)": ["vpf_gemm_bf16:atomic"],
    "colsum": ["vpf_colsum"],
    "bn_apply+bn_bwd": ["vpf_bn_apply", "vpf_bn_bwd"],
    "group_max_bwd+group_sum": ["vpf_group_max_bwd", "vpf_group_sum"],
    "linear3*": ["vpf_linear3_fwd", "vpf_linear3_stats", "vpf_linear3_bwd", "vpf_linear3_bn_bwd"],
    "tokenizer": ["vpf_divide_patches", "vpf_fps", "vpf_knn_group"],
    "ntxent+l2norm": ["vpf_ntxent_fwd", "vpf_ntxent_bwd", "vpf_l2norm_rows", "vpf_l2norm_bwd"],
    "adamw": ["vpf_adamw"],
    "add_scale+copy2d+cast": ["vpf_add_scale", "vpf_copy2d", "vpf_cast_bf16"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--config", default="A")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch

    import _synth
    import bench
    from vipformer_b200 import _lib, ops
    from vipformer_b200.engine import PretrainEngine

    bench.select_config(args.config)
    cfg = dict(bench.CFG, b=args.pairs)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    real_call = _lib.call
    skip = set()
    atomic_only = [False]

    def call(name, *a):
        if name in skip:
            return
        if atomic_only[0] and name == "vpf_gemm_bf16" and a[10]._obj.mode == ops.EPI_ATOMIC_ADD:
            return
        real_call(name, *a)

    _lib.call = call
    g = torch.Generator(device=dev).manual_seed(100)
    b = args.pairs
    pts = torch.randn((2 * b, cfg["N"], 3), device=dev, generator=g)
    img = torch.randn((b, 3, cfg["img"], cfg["img"]), device=dev, generator=g)
    base = None
    fams = {k: v for k, v in FAMILIES.items() if not args.only or k == "baseline" or k in args.only.split(",")}
    for fam, names in fams.items():
        skip.clear()
        atomic_only[0] = False
        for n in names:
            if n.endswith(":atomic"):
                atomic_only[0] = True
            else:
                skip.add(n)
        pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
        eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"], lr=1e-3, seed=1)
        eng.pc_in.copy_(pts); eng.img_in.copy_(img)
        for _ in range(4):
            eng.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            eng.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if base is None:
            base = ms
        print(f"{fam:28s} step {ms:7.3f} ms   marginal {base - ms:7.3f} ms", flush=True)
        eng.graph = None
        del eng, pc, im
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
