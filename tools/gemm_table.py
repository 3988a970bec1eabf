"""Per-GEMM efficiency table for one eager pre-training step (CUDA events around every launch)."""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth, bench
from vipformer_b200 import ops
from vipformer_b200.engine import PretrainEngine

b = 256
cfg = dict(bench.CFG, b=b)
pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)
eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"], use_cuda_graph=False, overlap_branches=False)
g = torch.Generator(device="cuda").manual_seed(1)
eng.pc_in.copy_(torch.randn(eng.pc_in.shape, device="cuda", generator=g) * 0.3)
eng.img_in.copy_(torch.randn(eng.img_in.shape, device="cuda", generator=g))
rec = []
orig = ops.gemm
def timed(a, b_, out, **kw):
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    M = a.shape[1] if a_mn else a.shape[0]; K = a.shape[0] if a_mn else a.shape[1]; N = b_.shape[1] if b_mn else b_.shape[0]
    mode = kw.get("mode", 0)
    osz = 0 if out is None else (4 if out.dtype == torch.float32 else 2)
    byt = 2 * M * K + 2 * N * K + M * N * osz + (M * N * 4 if kw.get("resid") is not None else 0) + (M * N * 2 if kw.get("aux") is not None else 0)
    if mode == 2: byt = 2 * M * K + 2 * N * K
    tag = f"{'T' if a_mn else 'N'}{'T' if b_mn else 'N'} M={M} N={N} K={K} mode={mode} out={'none' if out is None else str(out.dtype)[6:]}{' gm' if kw.get('gm_S') else ''}{' resid' if kw.get('resid') is not None else ''}{' aux' if kw.get('aux') is not None else ''}{' rg' if kw.get('rg_bias') is not None else ''}"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(a, b_, out, **kw); e1.record()
    rec.append((tag, e0, e1, 2.0 * M * N * K, byt))
    return r
ops.gemm = timed
for _ in range(2):
    rec.clear(); eng._step_body(); torch.cuda.synchronize()
ops.gemm = orig
agg = collections.OrderedDict()
for tag, e0, e1, fl, by in rec:
    a = agg.setdefault(tag, [0, 0.0, fl, by]); a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3
rows = []
for tag, (n, t, fl, by) in agg.items():
    ideal = max(fl / 1.371e15, by / 6.5427e12) * 1e6
    rows.append((t - n * ideal, tag, n, t / n, ideal, fl / (t / n) / 1e6, by / (t / n) / 1e3))
tot = sum(r[2] * r[3] for r in rows)
print(f"total gemm {tot/1e3:.2f} ms; ideal {sum(r[2]*r[4] for r in rows)/1e3:.2f} ms")
for lost, tag, n, t, ideal, tf, gb in sorted(rows, reverse=True)[:40]:
    print(f"lost {lost:7.0f}us  n={n:3d} avg {t:7.1f}us ideal {ideal:6.1f}us  {tf:6.0f} TF/s {gb:6.0f} GB/s  {tag}")
