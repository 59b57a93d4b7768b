set -u
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --env-envs-per-gpu 2048 --env-steps 120 --ddqn-envs-per-gpu 4096 --ddqn-steps 60"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:abx_dq_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_dq2 $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:abx_env_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_env2 $CMD > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -4
