"""Tokenizer micro-benchmark (device-resident inputs, CUDA-event timing)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth  # noqa: E402
from vipformer_b200.preproc import divide_patches, farthest_point_sample  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    G = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    pts = torch.from_numpy(_synth.make_clouds("randn", B, N, 1)).cuda()
    start = torch.from_numpy(_synth.make_start(B, N, 1)).cuda()
    t_all, t_all_min = timeit(lambda: divide_patches(pts, G, 32, start_idx=start))
    t_fps, _ = timeit(lambda: farthest_point_sample(pts, G, start_idx=start))
    bytes_per_cloud = N * 12 + G * 12 + G * 32 * 12
    gbs = B * bytes_per_cloud / (t_all * 1e-3) / 1e9
    print(json.dumps({"B": B, "N": N, "G": G, "ms_divide_patches": t_all, "ms_min": t_all_min, "ms_fps": t_fps,
                      "ms_knn_gather": t_all - t_fps, "clouds_per_s": B / (t_all * 1e-3),
                      "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]}))


if __name__ == "__main__":
    main()
