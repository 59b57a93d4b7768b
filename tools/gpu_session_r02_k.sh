#!/bin/bash
# round 2, session k: ncu evidence -- launch list of a shortened default bench, --set full captures of the headline run kernel and of the rmsc01 (SHAPE_P3) kernel
set -u
mkdir -p gpurun_out
LIST="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --env-steps 60 --ddqn-steps 30 --envs-per-gpu 4096 --env-envs-per-gpu 2368 --ddqn-envs-per-gpu 2368 --rmsc03-envs-per-gpu 2368 --mr-envs-per-gpu 2368"
timeout -s KILL 600 $LIST > gpurun_out/r02_list_plain.json 2> gpurun_out/r02_list_plain.err; echo "plain list rc=$?"
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_ncu.csv $LIST > gpurun_out/r02_list_ncu.log 2>&1; echo "ncu list rc=$?"
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay"
timeout -s KILL 300 $CMD > gpurun_out/r02_plain_k.log 2> gpurun_out/r02_plain_k.err && timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_run_k $CMD > gpurun_out/r02_ncu_k.log 2>&1; echo "ncu run rc=$?"
CMD1="python bench.py --envs-per-gpu 2368 --steps 1 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-ddqn --no-whole-day --no-marketreplay"
timeout -s KILL 300 $CMD1 > gpurun_out/r02_plain_k1.log 2> gpurun_out/r02_plain_k1.err && timeout -s KILL 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:abx_run_kernel<.*5>" -s 1 -c 1 -f -o gpurun_out/r02_prof_p3_k $CMD1 > gpurun_out/r02_ncu_k1.log 2>&1; echo "ncu p3 rc=$?"
cp marl_optimal_execution_b200/libabides_b200.so gpurun_out/r02_prof_k.so
ls -la gpurun_out | tail -12
