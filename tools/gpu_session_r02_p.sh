#!/bin/bash
# round 2, session p: two-rank bench (weak scaling check) and the reference arm
set -u
mkdir -p gpurun_out
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 bench rc=$?"; tail -2 gpurun_out/r02_bench_n2.err
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref bench rc=$?"; tail -2 gpurun_out/r02_bench_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n2.json", "gpurun_out/r02_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g n_gpus %s e2e %.4g" % (d["value"], d["n_gpus"], d["e2e"]["value"]), {k: ("%.4g" % d[k]["value"]) for k in ("env", "ddqn", "rmsc03", "rmsc01") if k in d and isinstance(d[k], dict) and "value" in d[k]})
    except Exception as e:
        print(f, "parse failed", e)
PY
