#!/bin/bash
# round 2, session l: rmsc02 (subscriptions) on the GPU -- new parity tests, then the rmsc01 / rmsc02 bench blocks alone
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "rmsc02 or rmsc01" > gpurun_out/r02_gpu_tests_l.log 2>&1; echo "rmsc pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_l.log
timeout -s KILL 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-ddqn --no-whole-day --no-marketreplay > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_l.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_l.json").read().strip().splitlines()[-1])
r = d.get("rmsc01") or {}
print({k: r[k] for k in ("value", "ms_per_run", "messages_per_env_run", "error_envs") if k in r})
print({k: v for k, v in (r.get("rmsc02") or {}).items() if k != "workload"})
PY
