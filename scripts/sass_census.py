"""SASS opcode census of libalg_b200.so per kernel: tcgen05 MMA (UTCHMMA, .2CTA = cta_group::2), TMA loads (UTMALDG), TMEM
loads / stores (LDTM / STTM), tcgen05.commit (UTCBAR), legacy tensor-core ops (HMMA -- must be 0).
    python scripts/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "alg_b200", "libalg_b200.so")], capture_output=True, text=True).stdout
OPS = ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM", "UTCATOM", "HMMA", "MUFU.EX2", "FFMA2")
per, cur, arch = collections.OrderedDict(), None, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)[:110]
        per[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                per[cur][o] += 1
                break
print("arch:", sorted(arch))
tot = collections.Counter()
print(f"{'kernel':110s} " + " ".join(f"{o:>12s}" for o in OPS))
for k, c in per.items():
    if any(c[o] for o in OPS if o not in ("FFMA2", "MUFU.EX2")):
        print(f"{k:110s} " + " ".join(f"{c[o]:12d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL (all ' + str(len(per)) + ' kernels)':110s} " + " ".join(f"{tot[o]:12d}" for o in OPS))
