set -u
OUT=gpurun_out; mkdir -p $OUT
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-svo --no-incremental > $OUT/r2_pipe_bench_${N}gpu.json 2> $OUT/r2_pipe_bench_${N}gpu.err
  echo "N=$N exit $?"; tail -c 200 $OUT/r2_pipe_bench_${N}gpu.err | tail -2
done
