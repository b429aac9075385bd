mkdir -p gpurun_out
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
python bench.py --steps 20 --warmup 5 --no-deposit --no-cpu-baseline > gpurun_out/bench_nodep.json 2> gpurun_out/bench_nodep.err; cat gpurun_out/bench_nodep.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_move -s 3 -c 2 -o gpurun_out/prof_move_r1a python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_import -s 3 -c 1 -o gpurun_out/prof_import_r1a python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_b.log
