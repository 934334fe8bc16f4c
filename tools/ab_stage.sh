#!/usr/bin/env bash
# A/B of tracer variants built by tools/build_variant.py: per-kernel CUDA-event times of the headline frame.
#   bash tools/ab_stage.sh <tag> <variant> [...]     ("default" = csrc/libvgi.so)
set -u
TAG=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
    if [ "$V" = default ]; then unset VGI_LIBVGI_PATH; else export VGI_LIBVGI_PATH=$PWD/tools/_dev/libvgi_$V.so; fi
    timeout 200 python tools/stage_times.py > "gpurun_out/${TAG}_${V}.log" 2>&1
    echo "== $V: $(grep -h 'k_trace' "gpurun_out/${TAG}_${V}.log" | tr -s ' ' | tr '\n' ';') $(grep -h 'diffuse mean' "gpurun_out/${TAG}_${V}.log")"
done
