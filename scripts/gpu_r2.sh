#!/bin/bash
# usage: bash scripts/gpu_r2.sh <tag> <what>...   (run under gpurun; outputs in gpurun_out/)
#   tests | bench | dipole | ncudipole | ncuall | sanitize | any shell command
mkdir -p gpurun_out
TAG=$1; shift
Q="--no-e2e --no-cpu-baseline --no-parity-gate"
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ;;
testsall) python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "passed|failed|FAILED|PASSED.*(lockstep|golden)|Error|error:|norm-relative|lock-step max" | tail -40 | tee gpurun_out/tests_$TAG.txt ;;
driver) ( cd /tmp && $OLDPWD/host/bin/ipic3d_b200 :U:32000000 2>&1 | grep -E "Throughput|took" ) | tee gpurun_out/driver_$TAG.txt ;;
bench) python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; python scripts/short.py < gpurun_out/bench_$TAG.json ;;
benchq) python bench.py --steps 20 --warmup 5 $Q > gpurun_out/benchq_$TAG.json 2> gpurun_out/benchq_$TAG.err; tail -3 gpurun_out/benchq_$TAG.err; python scripts/short.py < gpurun_out/benchq_$TAG.json ;;
refarm) python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/ref_$TAG.json 2> gpurun_out/ref_$TAG.err; tail -3 gpurun_out/ref_$TAG.err; cat gpurun_out/ref_$TAG.json ;;
dipole) python bench.py --workload dipole --steps 20 --warmup 12 $Q > gpurun_out/dipole_$TAG.json 2> gpurun_out/dipole_$TAG.err; tail -3 gpurun_out/dipole_$TAG.err; python scripts/short.py < gpurun_out/dipole_$TAG.json ;;
ncudipole)
  ncu --set full --clock-control none --import-source on -k regex:k_move -s 14 -c 1 -f -o gpurun_out/prof_dipole_move_$TAG python bench.py --workload dipole --steps 3 --warmup 12 $Q > gpurun_out/ncu_dm_$TAG.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_dipole_move_$TAG.ncu-rep --mix --top 25 > gpurun_out/${TAG}_dipole_move_summary.txt 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"k_import|k_sort_outbox" -s 28 -c 2 -f -o /tmp/prof_dipole_import python bench.py --workload dipole --steps 3 --warmup 12 $Q > gpurun_out/ncu_di_$TAG.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_dipole_import.ncu-rep --mix --top 12 > gpurun_out/${TAG}_dipole_sort_import_summary.txt 2>&1
  ;;
ncumain)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 $Q > gpurun_out/ncu_l_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_move -s 3 -c 1 -f -o gpurun_out/prof_move_$TAG python bench.py --steps 3 --warmup 3 $Q > gpurun_out/ncu_m_$TAG.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_move_$TAG.ncu-rep --mix --top 12 > gpurun_out/${TAG}_move_summary.txt 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_import -s 3 -c 1 -f -o /tmp/prof_import python bench.py --steps 3 --warmup 3 $Q > gpurun_out/ncu_i_$TAG.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_import.ncu-rep --mix --top 12 > gpurun_out/${TAG}_import_summary.txt 2>&1
  ;;
ncuall)
  for k in k_init_uniform k_gather_current k_boundaries k_solver k_cell_energy k_field_energy k_sum_pairs k_count_misplaced k_gather_aos k_scatter_aos k_check_species k_move_chunk k_face_count k_face_scan k_face_copy k_trace k_copy_heads; do
    ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/prof_${k} python scripts/probe_kernels.py > /tmp/ncu_${k}.log 2>&1
    python scripts/ncu_summary.py /tmp/prof_${k}.ncu-rep --mix --top 8 > gpurun_out/${TAG}_${k}_summary.txt 2>&1; rm -f /tmp/prof_${k}.ncu-rep
  done
  ncu --set full --clock-control none --import-source on -k regex:k_move -s 5 -c 1 -f -o /tmp/prof_k_move_interp python scripts/probe_kernels.py > /tmp/ncu_k_move_interp.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_k_move_interp.ncu-rep --mix --top 8 > gpurun_out/${TAG}_k_move_interp_summary.txt 2>&1
  ;;
dipvar)
  for v in "" "IPIC_WS_DIPOLE=1" "IPIC_LIB=allscale_ipic3d_b200/lib/var_dep6.so" "IPIC_LIB=allscale_ipic3d_b200/lib/var_dep4.so" "IPIC_STAGES=2" "IPIC_STAGES=3"; do
    echo -n "dipole [$v] : "; env $v python bench.py --workload dipole --steps 20 --warmup 12 $Q 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
  done ;;
medium) python bench.py --workload medium --steps 20 --warmup 5 > gpurun_out/medium_$TAG.json 2> gpurun_out/medium_$TAG.err; tail -3 gpurun_out/medium_$TAG.err; python scripts/short.py < gpurun_out/medium_$TAG.json ;;
ncurest)
  for k in k_solver_wavefront k_init_deck_positions k_init_normal_speed k_copy_heads; do
    timeout 60 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/prof_${k} python scripts/probe_kernels.py > /tmp/ncu_${k}.log 2>&1
    python scripts/ncu_summary.py /tmp/prof_${k}.ncu-rep --mix --top 8 > gpurun_out/${TAG}_${k}_summary.txt 2>&1; rm -f /tmp/prof_${k}.ncu-rep
  done ;;
sanitize)
  timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_$TAG.log
  timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_$TAG.log
  ;;
*) eval "$what" ;;
esac
done
