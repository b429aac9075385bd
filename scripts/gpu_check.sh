#!/bin/bash
# usage: bash scripts/gpu_check.sh <tag> [tests] [bench] [ncu]   (run under gpurun; outputs in gpurun_out/)
mkdir -p gpurun_out
TAG=$1; shift
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ;;
bench) python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | python scripts/short.py ;;
benchq) python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/benchq_$TAG.json 2> gpurun_out/benchq_$TAG.err; tail -2 gpurun_out/benchq_$TAG.err; cat gpurun_out/benchq_$TAG.json | python scripts/short.py ;;
nodep) python bench.py --steps 20 --warmup 5 --no-deposit --no-cpu-baseline --no-e2e > gpurun_out/bench_nodep_$TAG.json 2> gpurun_out/bench_nodep_$TAG.err; tail -2 gpurun_out/bench_nodep_$TAG.err; cat gpurun_out/bench_nodep_$TAG.json | python scripts/short.py ;;
ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_move -s 3 -c 1 -o gpurun_out/prof_move_$TAG python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_m_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_move -s 3 -c 1 -o gpurun_out/prof_move_nodep_$TAG python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-deposit > gpurun_out/ncu_mn_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_import -s 3 -c 1 -o gpurun_out/prof_import_$TAG python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_i_$TAG.log 2>&1
  ;;
*) eval "$what" ;;
esac
done
