#!/bin/bash
# round 2, session s: ncu --set full of the shipped run kernel (final tree) with source view, for the per-function footprint table of the optimisation log
set -u
mkdir -p gpurun_out
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay"
timeout -s KILL 300 $CMD > gpurun_out/r02_plain_s.log 2> gpurun_out/r02_plain_s.err && timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_run_s $CMD > gpurun_out/r02_ncu_s.log 2>&1; echo "ncu rc=$?"
