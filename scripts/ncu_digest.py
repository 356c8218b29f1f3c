"""Digest of an `ncu --set full --import-source on` report of one kernel: headline metrics + top stall sites.
Usage: python scripts/ncu_digest.py report.ncu-rep [n_top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
print("kernel:", d.get("Kernel Name", ("?",))[0][:100])
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for k in hdr:
    if any(k == w or k.endswith("." + w) for w in want):
        print(f"  {k} = {d[k][0]} {d[k][1]}")
for k in hdr:
    if "smsp__average_warp" in k and "per_issue_active" in k and "not_issued" not in k:
        v = float(d[k][0].replace(",", ""))
        if v > 0.05:
            print("  stall", k.split("issue_stalled_")[-1].split("_per")[0], round(v, 3))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h) and r[ix["# Samples"]].strip().isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    st = {k.replace("stall_", ""): int(r[ix[k]] or 0) for k in keys}
    st = {k: v for k, v in st.items() if v > 0.1 * int(r[ix["# Samples"]])}
    print(r[ix["Address"]][-5:], r[ix["# Samples"]].rjust(7), f'{100*int(r[ix["# Samples"]])/tot:5.1f}%', r[ix["Source"]][:64].ljust(64), st)
