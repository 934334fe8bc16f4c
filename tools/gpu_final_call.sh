#!/usr/bin/env bash
# The evidence call of a round (one B200, from the repo root under gpurun): GPU parity suite, smoke(), the bench line,
# the ncu launch list, the per-kernel DRAM traffic pass and full captures of the three dominant kernels — all of the
# SAME build. Every stage is time-boxed and independent. Output: gpurun_out/<tag>_*.
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p "$OUT"
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "$OUT/${TAG}_${name}.log" 2>&1; echo "   exit $? ($(tail -n 1 "$OUT/${TAG}_${name}.log" | cut -c1-160))"; }
run pytest_gpu 900 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 20 --warmup 5
grep -h '^{' "$OUT/${TAG}_bench.log" | tail -n 1 > "$OUT/${TAG}_bench.json"
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo --no-incremental
run traffic 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/${TAG}_traffic.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo --no-incremental
run ncu_trace 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_main|k_trace_specular' -s 8 -c 2 \
    -o "$OUT/${TAG}_trace" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo --no-incremental
run ncu_inject 600 ncu --set full --clock-control none --import-source on -k regex:'k_inject' -s 4 -c 1 \
    -o "$OUT/${TAG}_inject" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo --no-incremental
ls -la "$OUT" | grep "${TAG}_" | tail -n 20
