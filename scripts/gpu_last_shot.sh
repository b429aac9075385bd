#!/bin/bash
# One short GPU call (about 45 s of commands): the big-outbox import and the cluster field solver.
# Every step has its own timeout and writes its own file under gpurun_out/, so a cut-off call still leaves what ran.
mkdir -p gpurun_out
T0=$(( $(date +%s%N) / 1000000 ))
stamp() { echo "[$(( $(date +%s%N) / 1000000 - T0 )) ms] $1" | tee -a gpurun_out/r2y_times.txt; }
rm -f gpurun_out/r2y_times.txt
stamp start
timeout 14 python scripts/probe_big_import.py 8000000 B > gpurun_out/r2y_big_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/r2y_big_probe.txt
stamp "probe B done"
timeout 16 python -m pytest tests/test_gpu_parity.py -q -k "big_outbox or nonuniform or test_overflow_relayout or uniform_benchmark_bit_exact_1m" > gpurun_out/r2y_big_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2y_big_tests.txt
stamp "big tests done"
IPIC_SOLVER_MODE=cluster timeout 8 python -m pytest tests/test_gpu_parity.py -q -k "full_pic_step or field_solvers or relayout_keeps_fields" > gpurun_out/r2y_cluster_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2y_cluster_tests.txt
stamp "cluster tests done"
IPIC_SOLVER_MODE=cluster timeout 5 python scripts/probe_solver.py > gpurun_out/r2y_solver_cluster.txt 2>&1
stamp "cluster probe done"
timeout 5 python scripts/probe_solver.py > gpurun_out/r2y_solver_coop.txt 2>&1
stamp "coop probe done"
timeout 8 python scripts/probe_big_import.py 8000000 E > gpurun_out/r2y_big_probe_E.txt 2>&1; echo "rc=$?" >> gpurun_out/r2y_big_probe_E.txt
stamp "probe E done"
tail -n 6 gpurun_out/r2y_big_probe.txt gpurun_out/r2y_big_tests.txt gpurun_out/r2y_cluster_tests.txt gpurun_out/r2y_solver_cluster.txt gpurun_out/r2y_solver_coop.txt gpurun_out/r2y_big_probe_E.txt
