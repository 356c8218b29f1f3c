"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list.
Usage: python scripts/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("alg::", "")[:48]
    v = float(d["Metric Value"].replace(",", ""))
    u, m = d["Metric Unit"], d["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1.0))
        agg[name]["ms"] += v
        agg[name]["n"] += 1
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        agg[name]["bytes"] += v
tot = sum(a["ms"] for a in agg.values())
print(f"{'kernel':48s} {'n':>6s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s} {'DRAM GB/s':>10s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    bw = a["bytes"] / (a["ms"] / 1e3) / 1e9 if a.get("bytes") and a["ms"] else float("nan")
    print(f"{k:48s} {int(a['n']):6d} {a['ms']:10.2f} {a['ms'] / a['n']:9.3f} {100 * a['ms'] / tot:6.1f}% {bw:10.0f}")
print(f"{'total':48s} {'':6s} {tot:10.2f}")
