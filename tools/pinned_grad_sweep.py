"""One-off: worst per-parameter gradient error (pinned choices) vs batch size, linear and NT-Xent upstream."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _synth
from test_oracle_model_golden import oracle_run
from test_parity_pinned_gpu import run_product, pins_from_tap, compare_grads, relfro

torch.set_num_threads(os.cpu_count())
for base, bs in (("cfgA", (8, 16)), ("cfgB", (8,))):
    for b in bs:
        cfg = dict(_synth.MODEL_CASES["cfgA"], b=b, seed=60 + b)
        if base == "cfgB":
            cfg.update(D=384, H=6, MR=4)
        for linear in (True, False):
            t0 = time.time()
            r = run_product(cfg, linear=linear)
            pins = pins_from_tap(r["tap"], cfg)
            o = oracle_run(cfg, pins=pins, linear=linear)
            bad, worst = compare_grads(r, o, 5e-2)
            errs = sorted(bad, key=lambda x: -x[2] if isinstance(x[2], float) else 0)[:5]
            print(base, "b", b, "linear" if linear else "ntxent", "worst %.4f" % worst, "n_bad(>5e-2)", len(bad), errs,
                  "back", round(relfro(r["pc_back"], o["pc_back"]), 4), round(relfro(r["im_back"], o["im_back"]), 4),
                  "feats", round(relfro(r["pc_feats"], o["pc_feats"]), 4), round(relfro(r["im_feats"], o["im_feats"]), 4),
                  "t %.0fs" % (time.time() - t0), flush=True)
