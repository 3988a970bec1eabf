"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (instructions, samples)."""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, ie, isamp, agg = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
    elif r[0] not in ("", "Function Name") and ie is not None:
        try:
            n, s = int(r[ie] or 0), int(r[isamp] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0]), r[1].strip()[:110]), [0, 0])
        a[0] += n
        a[1] += s
tot, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("total warp instructions", tot, "samples", ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot * 100:5.1f}% inst {a[1] / max(ts, 1) * 100:5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
