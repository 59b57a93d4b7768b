#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_g.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_g.log
timeout 1200 python bench.py > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_g.json').readline())
print('lob %.4g e2e %.4g wholeday %.4g mr %.4g | rmsc03 %.4g pov %.4g | env %.4g e2e %.4g | ddqn %.4g e2e %.4g train %.4g | cpu %.4g issue_frac %s' % (d['value'], d['e2e']['value'], d['whole_day']['value'], d['marketreplay']['value'], d['rmsc03']['value'], d['rmsc03']['with_pov_execution_agent']['value'], d['env']['value'], d['env']['e2e']['value'], d['ddqn']['value'], d['ddqn']['e2e']['value'], d['ddqn']['training']['value'], d['cpu_baseline']['value'], d['roofline'].get('issue_frac')))
PY
timeout 900 python tools/train_ddqn.py --envs-per-gpu 2368 --episodes 9 --out gpurun_out/r02_ddqn_learning_curve.json > gpurun_out/r02_train.log 2>&1; echo "train rc=$?"; tail -12 gpurun_out/r02_train.log
CMD="python bench.py --envs-per-gpu 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-ddqn --no-whole-day --no-marketreplay"
$CMD > gpurun_out/r02_plain_b.log 2> gpurun_out/r02_plain_b.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 5 -c 1 -f -o gpurun_out/r02_prof_run_b $CMD > gpurun_out/r02_ncu_b.log 2>&1
cp marl_optimal_execution_b200/libabides_b200.so gpurun_out/r02_prof_run_b.so
