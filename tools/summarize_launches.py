"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(d["Metric Value"].replace(",", ""))
    if d.get("Metric Unit") in ("us", "usecond"): v *= 1e3
    if d.get("Metric Unit") in ("ms", "msecond"): v *= 1e6
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {t/n/1e3:8.1f} us  {k}")
