#!/bin/bash
# round 2, session m: full GPU suite + full bench after the OU-square / publish changes
set -u
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_m.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gpu_tests_m.log
timeout -s KILL 1200 python bench.py > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_m.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_m.json").read().strip().splitlines()[-1])
r1 = d.get("rmsc01") or {}
print("lob %.4g e2e %.4g wholeday %.4g mr %.4g | rmsc03 %.4g | rmsc01 %.4g rmsc02 %.4g | env %.4g e2e %.4g | ddqn %.4g e2e %.4g train %.4g | cpu %.4g issue_frac %s" % (
    d["value"], d["e2e"]["value"], d["whole_day"]["value"], d["marketreplay"]["value"], d["rmsc03"]["value"], r1.get("value", 0), (r1.get("rmsc02") or {}).get("value", 0),
    d["env"]["value"], d["env"]["e2e"]["value"], d["ddqn"]["value"], d["ddqn"]["e2e"]["value"], d["ddqn"]["training"]["value"], d["cpu_baseline"]["value"], d["roofline"].get("issue_frac")))
PY
