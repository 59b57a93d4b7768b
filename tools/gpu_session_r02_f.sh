#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_f.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gpu_tests_f.log
timeout 900 python bench.py > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f.json').readline())
print('lob %.4g e2e %.4g | rmsc03 %.4g pov %.4g | env %.4g e2e %.4g | ddqn %.4g e2e %.4g train %.4g | cpu %.4g' % (d['value'], d['e2e']['value'], d['rmsc03']['value'], d['rmsc03']['with_pov_execution_agent']['value'], d['env']['value'], d['env']['e2e']['value'], d['ddqn']['value'], d['ddqn']['e2e']['value'], d['ddqn']['training']['value'], d['cpu_baseline']['value']))
PY
