"""Per-kernel DRAM traffic of one bench frame from an ncu metrics pass, stored with the hash of the kernel sources it was
measured on (bench.py prints `traffic` only when that hash matches the library it is running).

On the GPU box (one GPU, never a bench value):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/<tag>_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo
  python tools/ncu_traffic.py gpurun_out/<tag>_traffic.csv gpurun_out/<tag>_ncu_traffic.json
then copy the json to profiles/r2_ncu_traffic.json. Per kernel: mean over its launches of read + write bytes."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    import bench
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        name = re.sub(r"[<(].*", "", r["Kernel Name"].replace("void ", ""))
        if not name.startswith("k_"):
            continue
        d = per.setdefault((name, r["ID"]), {})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(r.get("Metric Unit", ""), 1.0)
    kernels = {}
    for (name, _), m in per.items():
        k = kernels.setdefault(name, {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "duration_us_under_ncu": 0.0})
        k["launches"] += 1
        k["dram_bytes_read"] += m.get("dram__bytes_read.sum", 0.0)
        k["dram_bytes_write"] += m.get("dram__bytes_write.sum", 0.0)
        k["duration_us_under_ncu"] += m.get("gpu__time_duration.sum", 0.0) / 1e3
    for k in kernels.values():
        n = k["launches"]
        for f in ("dram_bytes_read", "dram_bytes_write", "duration_us_under_ncu"):
            k[f] /= n
        k["traffic"] = k["dram_bytes_read"] + k["dram_bytes_write"]
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over bench.py --steps 2 ({os.path.basename(src)}); "
                     "per kernel: mean per launch over all its launches (cold-cache, serialised)",
           "source_hash": bench.kernel_source_hash(), "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"wrote {dst}: {len(kernels)} kernels, source_hash {out['source_hash']}")


if __name__ == "__main__":
    main()
