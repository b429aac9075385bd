#!/bin/bash
# usage: bash scripts/gpu_scale.sh <tag> <N> [N ...]   (under gpurun --gpus max(N)); bench lines -> gpurun_out/scale_<tag>_<N>.json
mkdir -p gpurun_out
TAG=$1; shift
for N in "$@"; do
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${TAG}_$N.json 2> gpurun_out/scale_${TAG}_$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_${TAG}_$N.json 2> gpurun_out/scale_${TAG}_$N.err
  fi
  tail -3 gpurun_out/scale_${TAG}_$N.err
  tail -1 gpurun_out/scale_${TAG}_$N.json | python scripts/short.py
done
