#!/bin/bash
# Round-1 ncu evidence (run under gpurun, one GPU): launch list of a reduced bench.py run, then one --set full capture of each
# of the three step kernels.  The plain run of the same command must exit 0 first (B200_PROFILING.md).
set -u
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --env-envs-per-gpu 2048 --env-steps 120 --ddqn-envs-per-gpu 4096 --ddqn-steps 60"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abx_dq_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_dq $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abx_qnet_forward -s 40 -c 1 -f -o gpurun_out/prof_qnet $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abx_env_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_env $CMD > gpurun_out/ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/prof_run $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | tail -12
