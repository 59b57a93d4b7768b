#!/bin/bash
# A/B of run-kernel optimisation candidates (same sources, -D switches; same box, interleaved, twice).  build: here; run: on the GPU box.
#   tools/ab_opt_variants.sh build "name:-DFLAG ..." ...      tools/ab_opt_variants.sh run name ...
set -e
cd "$(dirname "$0")/.."
PKG=marl_optimal_execution_b200
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
if [ "$1" = build ]; then
  shift; mkdir -p build/ab
  for spec in "$@"; do v=${spec%%:*}; fl=${spec#*:}; [ "$fl" = "$spec" ] && fl=""; ( nvcc $FL $fl -Xptxas -v -o build/ab/opt_$v.so $PKG/csrc/abx_sim.cu $PKG/csrc/abx_qnet.cu > build/ab/$v.log 2>&1 ) & done
  wait; ls -la build/ab/*.so; exit 0
fi
shift
mkdir -p gpurun_out; : > gpurun_out/ab_opt.log
EXTRA=${AB_EXTRA:-"--no-ddqn --no-env --no-rmsc03 --no-rmsc01 --no-whole-day --no-marketreplay"}
for rep in 1 2; do for v in "$@"; do
  echo -n "$v: " | tee -a gpurun_out/ab_opt.log
  ABX_LIB_PATH=$PWD/build/ab/opt_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('lob msgs/s %.4g err %d' % (d['value'], d['detail']['error_envs']), ' '.join('%s %.4g' % (k, d[k]['value']) for k in ('rmsc03','rmsc01','env','ddqn') if k in d))" | tee -a gpurun_out/ab_opt.log
done; done
