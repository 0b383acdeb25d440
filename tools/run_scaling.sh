#!/bin/bash
# Multi-GPU bench lines on one box (run through `gpurun --gpus 8`):  bash tools/run_scaling.sh <tag> "2 4 8"
# One JSON line per N into gpurun_out/evidence/<tag>_bench_n<N>.json (the command the driver uses for N > 1).
set -u
TAG=${1:-r2}
NS=${2:-"2 8"}
O=gpurun_out/evidence
mkdir -p $O
port=29511
for n in $NS; do
  port=$((port + 1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n 2> $O/${TAG}_bench_n$n.err | tail -1 > $O/${TAG}_bench_n$n.json
  echo "== N=$n"; cut -c1-420 $O/${TAG}_bench_n$n.json; tail -2 $O/${TAG}_bench_n$n.err | cut -c1-300
done
