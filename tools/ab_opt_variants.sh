#!/bin/bash
# A/B of run-kernel optimisation candidates (same sources, -D switches; same box, interleaved, twice).  build: here; run: on the GPU box.
set -e
cd "$(dirname "$0")/.."
PKG=marl_optimal_execution_b200
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
declare -A V
V[noise4]="-DABX_NOISE4"
ORDER="noise4 cur"
if [ "$1" = build ]; then
  mkdir -p build/ab
  for v in $ORDER; do ( nvcc $FL ${V[$v]} -o build/ab/opt_$v.so $PKG/csrc/abx_sim.cu $PKG/csrc/abx_qnet.cu ) & done
  wait; ls -la build/ab; exit 0
fi
mkdir -p gpurun_out; : > gpurun_out/ab_opt.log
for rep in 1 2; do for v in $ORDER; do
  EXTRA="--no-ddqn --no-env --no-rmsc03 --no-rmsc01 --no-whole-day --no-marketreplay"
  echo -n "$v: " | tee -a gpurun_out/ab_opt.log
  ABX_LIB_PATH=$PWD/build/ab/opt_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('lob msgs/s %.4g err %d' % (d['value'], d['detail']['error_envs']), ' '.join('%s %.4g' % (k, d[k]['value']) for k in ('rmsc03','env','ddqn') if k in d))" | tee -a gpurun_out/ab_opt.log
done; done
