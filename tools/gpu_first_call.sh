#!/usr/bin/env bash
# First GPU call of a round, to be run under gpurun from the repo root (one B200):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh r2_v1'
# Runs the GPU parity suite, smoke(), the bench line, the launch list of a 2-step bench and one full ncu capture of the two
# tracer kernels and k_inject, everything into gpurun_out/<tag>_* (copy what is to be judged into profiles/ afterwards,
# summarised with tools/summarize_launches.py / tools/ncu_summary.py / tools/ncu_lines.py).
# Every stage is independent and time-boxed: a failing or hanging stage does not cost the others.
set -u
TAG=${1:-call}
OUT=gpurun_out
mkdir -p "$OUT"
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "$OUT/${TAG}_${name}.log" 2>&1; echo "   exit $? ($(tail -n 1 "$OUT/${TAG}_${name}.log" | cut -c1-160))"; }

nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > "$OUT/${TAG}_clocks.csv" 2>/dev/null &
SMI=$!

run pytest_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python __graft_entry__.py smoke
run bench 600 python bench.py --steps 20 --warmup 5
grep -h '^{' "$OUT/${TAG}_bench.log" | tail -n 1 > "$OUT/${TAG}_bench.json"

# launch list (cold-cache, serialised: shares, not absolutes); never a bench value
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo
# one full capture of the kernels the roofline objects name (3 launches each, after the warm-up launches)
run ncu_trace 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_main|k_trace_specular' -s 8 -c 4 \
    -o "$OUT/${TAG}_trace" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo
run ncu_inject 600 ncu --set full --clock-control none --import-source on -k regex:'k_inject' -s 4 -c 2 \
    -o "$OUT/${TAG}_inject" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo

kill $SMI 2>/dev/null
ls -la "$OUT" | tail -n 20
