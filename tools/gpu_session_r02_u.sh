#!/bin/bash
# round 2, session u: final default bench line (N=1) and the reference arm on the same box
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref bench rc=$?"; tail -2 gpurun_out/r02_bench_ref.err
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench_u.json 2> gpurun_out/r02_bench_u.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_u.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_u.json", "gpurun_out/r02_bench_ref.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), "issue_frac", d.get("roofline", {}).get("issue_frac"), {k: ("%.4g" % d[k]["value"]) for k in ("env", "ddqn", "rmsc03", "rmsc01", "whole_day", "marketreplay") if k in d and isinstance(d[k], dict) and "value" in d[k]})
PY
