"""Per-kernel-class time, DRAM traffic and achieved DRAM bandwidth from an ncu launch list that carries
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum (profiles/r01_step_launches_v5.csv)."""
import csv, json, os, re, sys
from collections import defaultdict

path = sys.argv[1]
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = None
for k, v in (peak.items() if isinstance(peak, dict) else []):
    if "hbm" in k.lower() and isinstance(v, (int, float)):
        hbm = float(v)
hbm = hbm or 6542.7
rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}
per = defaultdict(dict)
for r in rows:
    per[(r["ID"], r["Kernel Name"])][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, name), m in per.items():
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("vpf::", "")
    short = re.sub(r"<.*", "", short) if "gemm" not in short and "attn" not in short else short
    a = agg[short]
    a[0] += 1
    a[1] += m["gpu__time_duration.sum"][0] * T[m["gpu__time_duration.sum"][1]]
    a[2] += m["dram__bytes_read.sum"][0] * U[m["dram__bytes_read.sum"][1]]
    a[3] += m["dram__bytes_write.sum"][0] * U[m["dram__bytes_write.sum"][1]]
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | time (us) | share | DRAM read (MB) | DRAM written (MB) | achieved DRAM GB/s | of {hbm:.0f} GB/s |")
print("|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] < 0.002 * tot:
        continue
    bw = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9
    print(f"| `{k}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f} % | {a[2] / 1e6:.0f} | {a[3] / 1e6:.0f} | {bw:.0f} | {100 * bw / hbm:.0f} % |")
print(f"\ntotal {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
