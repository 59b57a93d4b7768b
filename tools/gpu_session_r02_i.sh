#!/bin/bash
# round 2, session i: the rmsc01 population (HBL / QUERY_ORDER_STREAM) on the GPU -- new parity tests first, then the whole suite and the bench
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "rmsc01" > gpurun_out/r02_gpu_tests_i_rmsc01.log 2>&1; echo "rmsc01 pytest rc=$?"; tail -12 gpurun_out/r02_gpu_tests_i_rmsc01.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_i.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_i.log
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_i.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_i.json").read().strip().splitlines()[-1])
    r1 = d.get("rmsc01") or {}
    print("lob %.4g e2e %.4g | rmsc03 %.4g | rmsc01 %.4g msgs/env %.0f err %s ms %.0f | env %.4g | ddqn %.4g" % (
        d["value"], d["e2e"]["value"], d["rmsc03"]["value"], r1.get("value", 0), r1.get("messages_per_env_run", 0), r1.get("error_envs"), r1.get("ms_per_run", 0),
        d["env"]["value"], d["ddqn"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
