#!/bin/bash
# round 2, session r: why do small code-size increases of the run kernel cost 5-18 %?  ncu --set full of the same launch for the shipped kernel and two A/B variants
set -u
mkdir -p gpurun_out
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay"
for v in ${AB_VARIANTS:-cur lvsign lanesKP}; do
  export ABX_LIB_PATH=$PWD/build/ab/opt_$v.so
  timeout -s KILL 300 $CMD > gpurun_out/r02_plain_r_$v.log 2> gpurun_out/r02_plain_r_$v.err && timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_r_$v $CMD > gpurun_out/r02_ncu_r_$v.log 2>&1; echo "ncu $v rc=$?"
done
ls -la gpurun_out | tail -8
