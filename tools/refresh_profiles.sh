#!/usr/bin/env bash
# Turn the outputs of tools/gpu_final_call.sh <tag> (gpurun_out/<tag>_*) into the tracked files under profiles/.
set -eu
T=$1; O=gpurun_out
cp $O/${T}_launches.csv profiles/r2_final_launches.csv
python tools/summarize_launches.py $O/${T}_launches.csv > profiles/r2_final_launches_summary.txt 2>&1
python tools/ncu_traffic.py $O/${T}_traffic.csv profiles/r2_ncu_traffic.json
cp $O/${T}_pytest_gpu.log profiles/r2_final_pytest_gpu.txt
cp $O/${T}_smoke.log profiles/r2_final_smoke.txt
python tools/ncu_summary.py $O/${T}_inject.ncu-rep 12 k_inject > profiles/r2_final_ncu_k_inject.txt 2>&1
python tools/ncu_summary.py $O/${T}_trace.ncu-rep 12 k_trace_main > profiles/r2_final_ncu_k_trace_main.txt 2>&1
python tools/ncu_summary.py $O/${T}_trace.ncu-rep 12 k_trace_specular_warp > profiles/r2_final_ncu_k_trace_specular_warp.txt 2>&1
