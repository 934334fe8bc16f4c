"""Aggregate tools/ncu_lines.py output by named source line ranges (functions).
Usage: ncu_regions.py report.ncu-rep object.o kernel  name:lo-hi [name:lo-hi ...]"""
import subprocess
import sys
import os
import re

rep, obj, kern = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep, obj, kern, "100000"],
                     capture_output=True, text=True).stdout
tot = {n: [0.0, 0.0] for n, _, _ in regions}
tot["other"] = [0.0, 0.0]
print(out.splitlines()[0])
for l in out.splitlines()[1:]:
    m = re.match(r"\s*(\d+|None)\s+([\d.]+)% inst\s+([\d.]+)% samples", l)
    if not m:
        continue
    ln = int(m.group(1)) if m.group(1) != "None" else -1
    for n, lo, hi in regions:
        if lo <= ln <= hi:
            tot[n][0] += float(m.group(2)); tot[n][1] += float(m.group(3))
            break
    else:
        tot["other"][0] += float(m.group(2)); tot["other"][1] += float(m.group(3))
for n, (a, b) in tot.items():
    print(f"{n:24s} {a:6.1f}% inst {b:6.1f}% samples")
