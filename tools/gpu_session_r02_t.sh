#!/bin/bash
# round 2, session t: final ncu evidence -- launch list of a shortened default bench and a --set full capture of the headline run kernel (final tree)
set -u
mkdir -p gpurun_out
LIST="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --env-steps 60 --ddqn-steps 30 --envs-per-gpu 4096 --env-envs-per-gpu 2368 --ddqn-envs-per-gpu 2368 --rmsc03-envs-per-gpu 2368 --mr-envs-per-gpu 2368"
timeout -s KILL 600 $LIST > gpurun_out/r02_list_plain.json 2> gpurun_out/r02_list_plain.err; echo "plain list rc=$?"
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_ncu.csv $LIST > gpurun_out/r02_list_ncu.log 2>&1; echo "ncu list rc=$?"
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay"
timeout -s KILL 300 $CMD > gpurun_out/r02_plain_t.log 2> gpurun_out/r02_plain_t.err && timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_run_t $CMD > gpurun_out/r02_ncu_t.log 2>&1; echo "ncu run rc=$?"
ls -la gpurun_out | tail -8
