#!/usr/bin/env bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient": nothing is charged).
# Usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   (extra gpurun flags via GPURUN_FLAGS, e.g. "--gpus 2")
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun ${GPURUN_FLAGS:-} --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
    rc=$?
    if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
    break
done
cat /tmp/gpurun_last.log
exit $rc
