"""Top stall sites of a kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:n]
for i in sorted(top):
    r = data[i]
    st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols if int(r[ix[h]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:90]:90s} {st}")
