#!/usr/bin/env bash
# Layout A/B of the voxel store as the tracer reads it (run under gpurun from the repo root, after
#   python tools/build_variant.py brick1 -DVGI_TRACE_BRICK_STORE=1; python tools/build_variant.py brick2 -DVGI_TRACE_BRICK_STORE=2):
# per variant the trace parity tests, the CUDA-event kernel times of a short bench, and one ncu metrics pass over the two tracer
# kernels (sectors per request, L1 / L2 hit rate, L1 and SM throughput, executed warp instructions).
set -u
TAG=${1:-layout}
OUT=gpurun_out
mkdir -p "$OUT"
bash tools/ab_bench.sh "$TAG" default brick1 brick2
M=gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_active
M=$M,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum
M=$M,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,dram__bytes_read.sum
for V in default brick1 brick2; do
    if [ "$V" = default ]; then unset VGI_LIBVGI_PATH; else export VGI_LIBVGI_PATH=$PWD/tools/_dev/libvgi_$V.so; fi
    timeout 600 ncu --metrics "$M" --clock-control none -k regex:'k_trace_main|k_trace_specular_warp' -s 8 -c 2 --csv \
        --log-file "$OUT/${TAG}_${V}_metrics.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-svo --no-incremental \
        > "$OUT/${TAG}_${V}_ncu.log" 2>&1
    echo "== $V ncu exit $?"
done
