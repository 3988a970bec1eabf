"""profiles/r02_step_dram.json + a per-kernel markdown table from the ncu launch list of one step
(`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... tools/profile_step.py`).

    python tools/dram_summary.py <launches.csv> <config A|B> <pairs> <out.json> [<out.md>]

bench.py quotes roofline.traffic from the JSON (only when config and pairs/GPU match the benchmarked ones)."""
import collections, csv, json, re, subprocess, sys

src, config, pairs, out_json = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out_md = sys.argv[5] if len(sys.argv) > 5 else None
rows = list(csv.reader(open(src, errors="ignore")))
hdr, launches = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("vpf::", "")
    L = launches.setdefault(d["ID"], {"name": name})
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "")
    if d["Metric Name"] == "gpu__time_duration.sum":
        L["us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        L["rd" if "read" in d["Metric Name"] else "wr"] = v * mult
agg = collections.OrderedDict()
for L in launches.values():
    a = agg.setdefault(L["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1; a["us"] += L.get("us", 0.0); a["rd"] += L.get("rd", 0.0); a["wr"] += L.get("wr", 0.0)
tot = {k: sum(a[k] for a in agg.values()) for k in ("n", "us", "rd", "wr")}
def cls(pred):
    sel = [a for k, a in agg.items() if pred(k)]
    n = sum(a["n"] for a in sel); b = sum(a["rd"] + a["wr"] for a in sel); us = sum(a["us"] for a in sel)
    return {"launches": n, "bytes": b, "bytes_per_launch": b / max(1, n), "us": us}
git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
js = {"git": git, "config": config, "pairs": pairs, "source": src, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
      "dram__bytes_write.sum --clock-control none, one eager serial step (tools/profile_step.py), cold caches per launch",
      "step": {"launches": tot["n"], "us": tot["us"], "dram_read_bytes": tot["rd"], "dram_write_bytes": tot["wr"],
               "dram_bytes": tot["rd"] + tot["wr"]},
      "gemm": cls(lambda k: k.startswith("gemm_bf16_kernel")), "attention": cls(lambda k: "attn" in k)}
json.dump(js, open(out_json, "w"), indent=1)
if out_md:
    with open(out_md, "w") as f:
        f.write(f"one step, config {config}, {pairs} pairs/GPU, git {git}: {tot['n']} launches, {tot['us'] / 1e3:.2f} ms serial, "
                f"{tot['rd'] / 1e9:.1f} GB read + {tot['wr'] / 1e9:.1f} GB written\n\n")
        f.write("| kernel | launches | time (us) | share | DRAM read (MB) | DRAM written (MB) | achieved DRAM GB/s |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"| `{k}` | {a['n']} | {a['us']:.0f} | {100 * a['us'] / tot['us']:.1f} % | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | "
                    f"{(a['rd'] + a['wr']) / max(a['us'], 1e-9) / 1e3:.0f} |\n")
print(json.dumps(js["step"]), json.dumps(js["gemm"]), json.dumps(js["attention"]))
