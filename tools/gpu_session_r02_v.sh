#!/bin/bash
# round 2, session v: ncu --set full of one mid-day ABIDESEnv step launch (abx_env_step_kernel<0,1>) of the final tree, with source view
set -u
mkdir -p gpurun_out
CMD="python bench.py --envs-per-gpu 2368 --steps 1 --warmup 3 --no-cpu-baseline --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay --env-steps 400"
timeout -s KILL 300 $CMD > gpurun_out/r02_plain_v.log 2> gpurun_out/r02_plain_v.err && timeout -s KILL 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:abx_env_step_kernel<.*1>" -s 300 -c 1 -f -o gpurun_out/r02_prof_env_v $CMD > gpurun_out/r02_ncu_v.log 2>&1; echo "ncu rc=$?"
