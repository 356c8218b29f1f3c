"""Executed warp instructions per CUDA source line of one kernel in an ncu report (needs -lineinfo + --import-source on).
Usage: python scripts/ncu_lines.py rep.ncu-rep <kernel regex> [n_top] [n-th matching launch, default 1]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
which = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # n-th matching launch
func, names = 0, []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, tot = {}, 0
hdr = None
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if r and r[0] == "Function Name":
        if r[1] not in names:
            names.append(r[1])
        func = names.index(r[1]) + 1
    if func != which:
        continue
    if hdr and r[0] and len(r) > hdr["Instructions Executed"] and r[hdr["Instructions Executed"]].isdigit():
        n = int(r[hdr["Instructions Executed"]])
        key = (fname, r[0])
        if key not in agg:
            agg[key] = [0, r[1].strip()[:100]]
        agg[key][0] += n
        tot += n
print("total warp instructions", tot)
items = [(v[0], k[0], k[1], v[1]) for k, v in agg.items()]
for n, f, ln, src in sorted(items, reverse=True)[:ntop]:
    print(f"{100 * n / tot:5.1f}%  {f}:{ln:>4s}  {src}")
