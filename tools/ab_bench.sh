#!/usr/bin/env bash
# A/B of libvgi variants built by tools/build_variant.py (run under gpurun from the repo root):
#   bash tools/ab_bench.sh <tag> <variant> [<variant> ...]       ("default" = the in-tree csrc/libvgi.so)
# For each variant: the trace parity tests (oracle comparison) and a short bench; prints the per-kernel CUDA-event times.
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p "$OUT"
for V in "$@"; do
    if [ "$V" = default ]; then unset VGI_LIBVGI_PATH; else export VGI_LIBVGI_PATH=$PWD/tools/_dev/libvgi_$V.so; fi
    echo "== $V"
    timeout 600 python -m pytest tests/test_gpu_clipmap.py tests/test_gpu_fullsize.py -m gpu -x -q -k "trace or headline or mode or cone" > "$OUT/${TAG}_${V}_pytest.log" 2>&1
    echo "   pytest exit $? ($(tail -n 1 "$OUT/${TAG}_${V}_pytest.log" | cut -c1-120))"
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-svo > "$OUT/${TAG}_${V}_bench.log" 2>&1
    grep -h '^{' "$OUT/${TAG}_${V}_bench.log" | tail -n 1 > "$OUT/${TAG}_${V}_bench.json"
    python - "$OUT/${TAG}_${V}_bench.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    k = d.get("kernels", {})
    print("   value", round(d["value"], 1), "fps; trace_ms", round(d["stages"]["trace_ms"], 3), "build_ms", round(d["stages"]["build_ms"], 3))
    print("   " + "  ".join(f"{n}={v['ms_per_step']:.3f}" for n, v in k.items() if v["ms_per_step"] > 0.02))
except Exception as e:
    print("   no bench line:", e)
PY
done
