#!/usr/bin/env bash
# 2 / 4 / 8-GPU bench lines exactly as the driver launches them (run under `gpurun --gpus 8` from the repo root).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2_scale}
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
  echo "N=$N exit $?"
done
