#!/usr/bin/env python
"""bench.py -- ViPFormer pre-training step throughput on B200 (BASELINE.json metric: shapes/sec, pretrain step
E1CL8SL-H4D256-L128-MR2; 1 shape = 1 pair = 2 point-cloud views + 1 image).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Keys as the driver's contract: value = whole-job shapes/s with inputs resident in HBM
(CUDA-event timed, max over ranks); e2e = the same metric through the public API with pinned HOST buffers (H2D + D2H
inside the timed region); roofline = the dominant kernel (tcgen05 GEMM) measured live with CUDA events;
cpu_baseline = the oracle port (plain PyTorch fp32 restatement of the reference) timed on this box's host cores.
`--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs[1] (the metric's configuration, default) and configs[2]; FLOPs: BASELINE.md section 2 (3 x forward)
CFGS = {
    "A": dict(name="E1CL8SL-H4D256-L128-MR2", flop=24.76e9, cfg=dict(D=256, H=4, n_sa=8, G=128, S=32, N=2048, MR=2, img=144, patch=12, seed=1)),
    "B": dict(name="E1CL8SL-H6D384-L128-MR4", flop=58.81e9, cfg=dict(D=384, H=6, n_sa=8, G=128, S=32, N=2048, MR=4, img=144, patch=12, seed=1)),
}
CFG = CFGS["A"]["cfg"]
WORKLOAD = "pretrain step E1CL8SL-H4D256-L128-MR2: fwd (pc 2x + img) + NT-Xent (intra+cross) + bwd + grad all-reduce + AdamW"
FLOP_PER_SHAPE_STEP = 24.76e9


def select_config(key):
    global CFG, WORKLOAD, FLOP_PER_SHAPE_STEP
    c = CFGS[key]
    CFG, FLOP_PER_SHAPE_STEP = c["cfg"], c["flop"]
    WORKLOAD = f"pretrain step {c['name']}: fwd (pc 2x + img) + NT-Xent (intra+cross) + bwd + grad all-reduce + AdamW"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_rate(pairs, steps, warmup, threads=None):
    """The oracle port (oracle/model_ref.py: plain PyTorch fp32 restatement of the reference modules + restated
    NT-Xent) running the same training step on the host cores: fwd + loss + bwd + torch.optim.AdamW."""
    import numpy as np
    import torch

    import _synth
    from oracle import model_ref as M

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(CFG, b=pairs)
    pc, im = _synth.build_models(cfg)
    sd_pc = {k: v.clone() for k, v in pc.state_dict().items()}
    sd_im = {k: v.clone() for k, v in im.state_dict().items()}
    plist = []
    for sd, model in ((sd_pc, pc), (sd_im, im)):
        for k, _ in model.named_parameters():
            sd[k] = sd[k].requires_grad_(True)
            plist.append(sd[k])
        for k in list(sd):
            if "cross_attn_n." in k:
                sd[k.replace("cross_attn_n.", "cross_attn_1.")] = sd[k]
    opt = torch.optim.AdamW(plist, lr=1e-3)
    pts, start, imgs = _synth.model_inputs(cfg)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        pf, _ = M.pc_forward(sd_pc, pts, start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], True)
        jf, _ = M.img_forward(sd_im, imgs, cfg["patch"], cfg["H"], cfg["n_sa"], True)
        total, _, _ = M.pretrain_loss(pf, jf)
        total.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(times))
    return pairs / (ms * 1e-3), ms, threads


def gpu_eager_step_rate(pairs, steps, warmup, dev):
    """The stronger comparator of SURVEY.md 8(d): the reference's math in STOCK PyTorch eager on the same B200 -- the
    oracle port (oracle/model_ref.py + oracle/tokenizer_torch.py: the ATen op sequence the reference itself executes)
    under bf16 autocast, cuBLAS / ATen kernels only, torch.optim.AdamW; same step, same batch, dropout off."""
    import numpy as np
    import torch

    import _synth
    from oracle import model_ref as M
    from oracle import tokenizer_torch as TT

    cfg = dict(CFG, b=pairs)
    pc, im = _synth.build_models(cfg)
    sd_pc = {k: v.clone().to(dev) for k, v in pc.state_dict().items()}
    sd_im = {k: v.clone().to(dev) for k, v in im.state_dict().items()}
    plist = []
    for sd, model in ((sd_pc, pc), (sd_im, im)):
        for k, _ in model.named_parameters():
            sd[k] = sd[k].requires_grad_(True)
            plist.append(sd[k])
        for k in list(sd):
            if "cross_attn_n." in k:
                sd[k.replace("cross_attn_n.", "cross_attn_1.")] = sd[k]
    del pc, im
    opt = torch.optim.AdamW(plist, lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(7)
    pts = torch.randn((2 * pairs, cfg["N"], 3), device=dev, generator=g)
    pts = pts - pts.mean(1, keepdim=True)
    pts = pts / pts.norm(dim=-1).amax(1).view(-1, 1, 1)
    imgs = torch.randn((pairs, cfg["img"], cfg["img"], 3), device=dev, generator=g)
    start = torch.randint(0, cfg["N"], (2 * pairs,), device=dev, generator=g)
    tok = lambda p, G, S, st: TT.divide_patches(p, G, S, st, sorted_knn=False)
    times = []
    for it in range(warmup + steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pf, _ = M.pc_forward(sd_pc, pts, start, cfg["G"], cfg["S"], cfg["H"], cfg["n_sa"], True, tokenizer=tok)
            jf, _ = M.img_forward(sd_im, imgs, cfg["patch"], cfg["H"], cfg["n_sa"], True)
        total, _, _ = M.pretrain_loss(pf.float(), jf.float())
        total.backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    del opt, plist, sd_pc, sd_im
    torch.cuda.empty_cache()
    return pairs / (ms * 1e-3), ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = 16
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    rate, ms, threads = cpu_reference_step_rate(pairs, steps, warmup)
    sample = f"{pairs} pairs/step x {steps} steps (+{warmup} warm-up), fp32, torch {threads} threads, dropout off"
    print(json.dumps({
        "impl": "reference", "metric": "shapes/sec", "value": rate, "unit": "shapes/s", "n_gpus": 0, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "pairs_per_step": pairs},
        "cpu_baseline": {"value": rate, "unit": "shapes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import _synth
    from vipformer_b200 import _lib, ops
    from vipformer_b200.engine import PretrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: vipformer_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    b = args.pairs
    cfg = dict(CFG, b=b)
    pc, im = _synth.build_models(cfg, atten_drop=0.1, mlp_drop=0.5)     # script values (scripts/pretrain/*.sh)
    eng = PretrainEngine(pc, im, batch_pairs=b, num_points=cfg["N"], img_size=cfg["img"], lr=1e-3, seed=1,
                         use_cuda_graph=not args.no_graph,
                         overlap_branches=not args.no_overlap)
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def synth_clouds(n):
        p = torch.randn((n, cfg["N"], 3), device=dev, generator=g)
        p = p - p.mean(1, keepdim=True)
        return p / p.norm(dim=-1).amax(1).view(n, 1, 1)

    # a few distinct device-resident batches so consecutive steps do not re-read identical inputs
    batches = [(synth_clouds(b), synth_clouds(b), torch.randn((b, 3, cfg["img"], cfg["img"]), device=dev, generator=g))
               for _ in range(2)]
    host = [tuple(t.cpu().pin_memory() for t in bt) for bt in batches]

    def load(i):
        t1, t2, im_ = batches[i % len(batches)]
        eng.pc_in[:b].copy_(t1); eng.pc_in[b:].copy_(t2); eng.img_in.copy_(im_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(msg):
        if args.verbose and rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    # ---- warm-up (includes graph capture) + launch accounting
    _lib.launch_count_reset()
    load(0)
    eng.step()
    torch.cuda.synchronize()
    launches_capture = _lib.launch_count()
    phase("first step (capture) done")
    launches_per_step = launches_capture // 3 if eng.graph is not None else launches_capture
    for i in range(max(args.warmup, 3)):
        load(i)
        eng.step()
    barrier()

    # ---- timed region: device-resident inputs, CUDA events on the launching stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        load(i)
        eng.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    loss_val = eng.losses.cpu().tolist()
    phase(f"timed region done: {ms / args.steps:.2f} ms/step")
    # ---- e2e: pinned host -> device every step, loss read back every step
    for i in range(min(2, args.warmup)):      # untimed: creates the copy stream / staging buffers of the host path
        if args.e2e_sync:
            eng.step_host(*host[i % len(host)])
        else:
            eng.prefetch_host(*host[i % len(host)])
            eng.step_host_prefetched(None)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if args.e2e_sync:
        for i in range(args.steps):
            eng.step_host(*host[i % len(host)])
    else:
        # input pipeline as a DataLoader with pinned memory + prefetch provides it: the H2D copy of step i+1 runs on a
        # copy stream while step i computes; every step's inputs are copied (and its losses read back) inside the timed region
        # The losses of step i are read on the host while step i+1 runs (one-step lag, as a loop that logs the
        # previous iteration's loss.item() does); the last step's are read before the clock stops.
        eng.prefetch_host(*host[0])
        for i in range(args.steps):
            eng.step_host_prefetched(host[(i + 1) % len(host)] if i + 1 < args.steps else None, lag_losses=True)
        last = eng.drain_losses()
        assert last is not None and bool(torch.isfinite(last).all())
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    phase("e2e done")
    clocks = sampler.stop() if rank == 0 else None

    tms = torch.tensor([ms, max(ms_e2e, wall_e2e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tms.tolist()
    value = world * b * args.steps / (ms * 1e-3)
    e2e_value = world * b * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = eng.losses_host.numel() * 4

    out = None
    pk, pk_kind = peaks()
    # roofline of the dominant kernel (tcgen05 GEMM): eager steps with CUDA events around every GEMM launch.  The step
    # contains collectives when world > 1, so EVERY rank runs it; rank 0 reports.
    roof, launches_step = gemm_roofline(eng, load, pk, pk_kind, args)
    phase("roofline pass done")
    barrier()
    if rank == 0:
        tok = tokenizer_rate(dev, pk)
        # the CPU arm is timed on rank 0 at N = 1 only (it would only delay the other ranks' exit at N > 1)
        cpu_rate, cpu_ms, threads = cpu_reference_step_rate(16, 3, 1) if world == 1 else (None, None, None)
        eager = None
        if world == 1 and not args.no_eager_baseline:
            torch.cuda.empty_cache()
            try:
                er, ems = gpu_eager_step_rate(b, 3, 2, dev)
                eager = {"value": er, "unit": "shapes/s", "ms_per_step": ems, "ratio_ours_over_eager": value / er,
                         "what": "oracle port (reference math, ATen/cuBLAS kernels) in stock PyTorch eager under bf16 autocast + "
                                 "torch.optim.AdamW on this GPU, same config and pairs/step, dropout off, 3 steps after 2 warm-up"}
            except Exception as ex:      # e.g. out of memory at a large --pairs: report, do not fail the bench line
                eager = {"value": None, "error": f"{type(ex).__name__}: {str(ex)[:120]}"}
        phase("baselines done")
        out = {
            "metric": "shapes/sec", "value": value, "unit": "shapes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "config_key": args.config, "pairs_per_gpu": b, "global_pairs": b * world, "points": cfg["N"],
                       "parallelism": f"dp{world}", "negatives": "global (all-gather)" if eng.gather else "rank-local",
                       "two_stream_branches": eng.side is not None, "e2e_input": "synchronous" if args.e2e_sync else "H2D of step i+1 prefetched on a copy stream during step i; losses of step i read back (D2H) while step i+1 runs, the last before the clock stops", "cuda_graph": eng.graph is not None, "dropout": "atten 0.1 / mlp 0.5",
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; 2 alternating input batches"},
            "e2e": {"value": e2e_value, "unit": "shapes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_step * args.steps),
            "gpu_launches_per_step": int(launches_step),
            "clocks": clocks,
            "roofline": roof,
            "model_flops_fraction": {"flop_per_shape_step": FLOP_PER_SHAPE_STEP,
                                     "achieved_tflops_per_gpu": value / world * FLOP_PER_SHAPE_STEP / 1e12,
                                     "peak_tflops": pk["bf16_tflops_sustained"], "peak_kind": pk_kind + " sustained",
                                     "frac": value / world * FLOP_PER_SHAPE_STEP / 1e12 / pk["bf16_tflops_sustained"]},
            "tokenizer": tok,
            "gpu_eager_baseline": eager,
            "cpu_baseline": {"value": cpu_rate, "unit": "shapes/s", "cores": threads, "kind": "port",
                             "sample": "16 pairs/step x 3 steps (+1 warm-up) of the same workload, oracle port fp32, dropout off"
                             if world == 1 else "measured at N = 1 only"},
            "loss": loss_val,
        }
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        # tear-down order matters with NCCL kernels captured in a CUDA graph: drop the graph before the communicator
        phase("teardown")
        dist.barrier()
        torch.cuda.synchronize()
        eng.close()
        del eng
        torch.cuda.synchronize()
        dist.destroy_process_group()


def gemm_roofline(eng, load, pk, pk_kind, args):
    """Run steps eagerly with a CUDA-event pair around every tcgen05 GEMM launch (same stream): achieved TFLOP/s =
    algorithmic FLOPs (2*M*N*K of each launch) / summed launch durations.  Also counts the kernels this library
    launches in one step (vpf_launch_count), and the algorithmic bytes of the GEMM launches (operands + result once)."""
    import torch

    from vipformer_b200 import _lib, ops

    rec = []
    orig = ops.gemm

    def timed_gemm(a, b_, out, **kw):
        a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
        M = kw.get("M") or (a.shape[1] if a_mn else a.shape[0])
        K = kw.get("K") or (a.shape[0] if a_mn else a.shape[1])
        N = kw.get("N") or (b_.shape[1] if b_mn else b_.shape[0])
        obytes = 0 if out is None else M * N * out.element_size() * (2 if kw.get("mode", 0) == 1 else 1)   # residual: read + write
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(a, b_, out, **kw)
        e1.record()
        rec.append((e0, e1, 2.0 * M * N * K, 2.0 * (M * K + N * K) + obytes))
        return r

    ops.gemm = timed_gemm
    side, eng.side = eng.side, None     # serial schedule: a launch timed while the other branch runs is not its own time
    launches = 0
    try:
        for i in range(2):
            rec.clear()
            load(i)
            torch.cuda.synchronize()
            c0 = _lib.launch_count()
            eng._step_body()
            torch.cuda.synchronize()
            launches = _lib.launch_count() - c0
    finally:
        ops.gemm = orig
        eng.side = side
    t = sum(e0.elapsed_time(e1) for e0, e1, _, _ in rec) * 1e-3
    fl = sum(f for _, _, f, _ in rec)
    ab = sum(b for _, _, _, b in rec)
    ach = fl / t / 1e12
    peak = pk["bf16_tflops_sustained"]
    # measured DRAM traffic of the GEMM launches: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu
    # launch list of `tools/profile_step.py` committed under profiles/ (tools/dram_summary.py writes the JSON); only
    # quoted when it was taken on the configuration being benchmarked
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "r02_step_dram.json")
    if os.path.exists(tj):
        d = json.load(open(tj))
        if d.get("pairs") == eng.b and d.get("config") == args.config:
            traffic = d["gemm"]["bytes_per_launch"]
            traffic_src = f"profiles/r02_step_dram.json (ncu, git {d.get('git', '?')}, {d['gemm']['launches']} launches)"
    # which roof binds?  arithmetic intensity of the step's GEMMs against the machine's ridge point
    hbm_peak = pk["hbm_gbs"]
    ai, ridge = fl / ab, peak * 1e12 / (hbm_peak * 1e9)
    ach_gbs = ab / t / 1e9
    frac_tensor, frac_hbm = ach / peak, ach_gbs / hbm_peak
    if ai < ridge:
        bound, achieved, pk_val, unit, frac = "hbm", ach_gbs, hbm_peak, "GB/s", frac_hbm
    else:
        bound, achieved, pk_val, unit, frac = "tensor", ach, peak, "TFLOP/s", frac_tensor
    roof = {"kernel": "gemm_bf16_kernel (tcgen05.mma + TMA)", "bound": bound, "achieved": achieved, "peak": pk_val,
            "unit": unit, "frac": frac,
            "bound_rule": f"arithmetic intensity {ai:.0f} FLOP/B {'<' if ai < ridge else '>='} ridge {ridge:.0f} FLOP/B "
                          "(sustained bf16 peak / measured copy bandwidth)",
            "frac_tensor": frac_tensor, "achieved_tflops": ach, "peak_tflops": peak,
            "frac_hbm": frac_hbm, "achieved_gbs": ach_gbs, "peak_gbs": hbm_peak,
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
            "traffic_source": traffic_src, "algorithmic_bytes": ab / max(1, len(rec)),
            "algorithmic_bytes_unit": "bytes/launch (operands + result once; residual read)",
            "peak_kind": pk_kind + " (sustained bf16 matmul; copy bandwidth) -- kernel timed inside a long step",
            "launches_per_step": len(rec), "avg_launch_us": 1e6 * t / max(1, len(rec)), "gemm_ms_per_step": 1e3 * t,
            "gemm_flops_per_step": fl}
    return roof, launches


def tokenizer_rate(dev, pk):
    import numpy as np
    import torch

    from vipformer_b200.preproc import divide_patches

    B, N, G = 512, 2048, 128
    g = torch.Generator(device=dev).manual_seed(5)
    pts = torch.randn((B, N, 3), device=dev, generator=g)
    start = torch.randint(0, N, (B,), device=dev, generator=g)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        divide_patches(pts, G, 32, start_idx=start)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    bytes_per_cloud = N * 12 + G * 12 + G * 32 * 12
    gbs = B * bytes_per_cloud / (ms * 1e-3) / 1e9
    return {"clouds_per_s": B / (ms * 1e-3), "ms_512_clouds": ms, "bound": "hbm (nominal); fp32-issue in practice",
            "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "algorithmic_bytes_per_cloud": bytes_per_cloud, "l2_flush": "256 MiB write between iterations"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU (weak scaling); the reference's own per-rank batch is 55")
    ap.add_argument("--config", default="A", choices=sorted(CFGS), help="A = E1CL8SL-H4D256-L128-MR2 (the metric's config), B = E1CL8SL-H6D384-L128-MR4")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the stock-PyTorch-eager comparator leg (N = 1)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-sync", action="store_true", help="e2e without input prefetch: H2D copy, step, D2H read in series")
    ap.add_argument("--no-overlap", action="store_true", help="run the image branch on the main stream (no two-stream overlap)")
    ap.add_argument("--verbose", action="store_true", help="progress lines on stderr (rank 0)")
    ap.add_argument("--graph", action="store_true", help="(kept for compatibility: CUDA-graph replay is the default at every world size)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
