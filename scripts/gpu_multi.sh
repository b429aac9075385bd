#!/bin/bash
# usage: bash scripts/gpu_multi.sh <tag> <N> <what>...   (run under gpurun --gpus N; outputs in gpurun_out/)
mkdir -p gpurun_out
TAG=$1; N=$2; shift; shift
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for what in "$@"; do
case $what in
nccltest) timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 ;;
ncclall) $TR tests/mgpu_nccl_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12 ;;
uniform) $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${TAG}_$N.json 2> gpurun_out/scale_${TAG}_$N.err; tail -3 gpurun_out/scale_${TAG}_$N.err; python scripts/short.py < gpurun_out/scale_${TAG}_$N.json ;;
uniformq) $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-parity-gate > gpurun_out/scaleq_${TAG}_$N.json 2> gpurun_out/scaleq_${TAG}_$N.err; tail -3 gpurun_out/scaleq_${TAG}_$N.err; python scripts/short.py < gpurun_out/scaleq_${TAG}_$N.json ;;
medium) $TR bench.py --gpus $N --workload medium --steps 20 --warmup 5 --no-e2e > gpurun_out/medium_${TAG}_$N.json 2> gpurun_out/medium_${TAG}_$N.err; tail -3 gpurun_out/medium_${TAG}_$N.err; python scripts/short.py < gpurun_out/medium_${TAG}_$N.json ;;
mediumq) $TR bench.py --gpus $N --workload medium --steps 20 --warmup 5 --no-e2e --no-parity-gate --device-init > gpurun_out/mediumq_${TAG}_$N.json 2> gpurun_out/mediumq_${TAG}_$N.err; tail -3 gpurun_out/mediumq_${TAG}_$N.err; python scripts/short.py < gpurun_out/mediumq_${TAG}_$N.json ;;
large) $TR bench.py --gpus $N --workload large --steps 5 --warmup 2 > gpurun_out/large_${TAG}_$N.json 2> gpurun_out/large_${TAG}_$N.err; tail -5 gpurun_out/large_${TAG}_$N.err; python scripts/short.py < gpurun_out/large_${TAG}_$N.json ;;
ref) python bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/ref_${TAG}_$N.json 2> gpurun_out/ref_${TAG}_$N.err; tail -2 gpurun_out/ref_${TAG}_$N.err; python scripts/short.py < gpurun_out/ref_${TAG}_$N.json ;;
*) eval "$what" ;;
esac
done
