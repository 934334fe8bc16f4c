"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
and mean device time, share of all launches. Usage: summarize_launches.py launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        rows.append((int(r["ID"]), name, float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {total/1e3:.1f} us total (ncu per-launch times are cold-cache, serialised: compare shares)")
    print(f"{'kernel':48s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share':>7s}  grid / block (last)")
    for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:48]:48s} {n:8d} {ns/1e3:10.1f} {ns/1e3/n:9.2f} {100*ns/total:6.1f}%  {grid} / {block}")


if __name__ == "__main__":
    main()
