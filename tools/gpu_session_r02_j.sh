#!/bin/bash
# round 2, session j: rmsc01 after the histogram form of the HBL belief table -- parity, then the rmsc01 bench block alone
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "rmsc01" > gpurun_out/r02_gpu_tests_j_rmsc01.log 2>&1; echo "rmsc01 pytest rc=$?"; tail -5 gpurun_out/r02_gpu_tests_j_rmsc01.log
timeout -s KILL 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-env --no-rmsc03 --no-ddqn --no-whole-day --no-marketreplay > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_j.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_j.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("rmsc01"), indent=1))
PY
