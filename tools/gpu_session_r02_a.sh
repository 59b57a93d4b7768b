#!/bin/bash
# Round-2 GPU session A: full GPU suite, sync-variant A/B, compute-sanitizer logs, ncu capture of the run kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gpu_tests_a.log
timeout 900 tools/ab_sync_variants.sh run; cat gpurun_out/ab_sync.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_parity.py > gpurun_out/r02_racecheck_default.log 2>&1; echo "racecheck default rc=$?"; tail -4 gpurun_out/r02_racecheck_default.log
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_parity.py > gpurun_out/r02_synccheck_default.log 2>&1; echo "synccheck default rc=$?"; tail -3 gpurun_out/r02_synccheck_default.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_parity.py marl_optimal_execution_b200/libabides_b200_strict.so > gpurun_out/r02_racecheck_strict.log 2>&1; echo "racecheck strict rc=$?"; tail -4 gpurun_out/r02_racecheck_strict.log
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-ddqn"
$CMD > gpurun_out/r02_plain_a.log 2> gpurun_out/r02_plain_a.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_run_a $CMD > gpurun_out/r02_ncu_a.log 2>&1
cp marl_optimal_execution_b200/libabides_b200.so gpurun_out/r02_prof_run_a.so
ls -la gpurun_out | tail -15
