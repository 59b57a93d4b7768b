#!/bin/bash
# round 2, session q: full GPU suite, smoke and the default bench line on the tree of the day
set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_q.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_gpu_tests_q.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_q.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_q.log
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_q.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_q.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g issue_frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["issue_frac"]), {k: ("%.4g" % d[k]["value"]) for k in ("env", "ddqn", "rmsc03", "rmsc01", "whole_day", "marketreplay") if k in d and isinstance(d[k], dict) and "value" in d[k]})
PY
