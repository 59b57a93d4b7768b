#!/bin/bash
# A/B of the on-chip synchronisation variants of the run / step kernels (same sources, same box, back to back):
#   default      xsync() = __syncwarp() at cross-lane handoffs, sync() = compiler fence between uniform stores and reads
#   strict       -DABX_STRICT_SYNC: every sync() a __syncwarp() as well
#   fence_only   -DABX_FENCE_ONLY: round-1 behaviour (no real barrier at the handoffs) -- measurement only, never shipped
# Build here (CPU): tools/ab_sync_variants.sh build ; run on the GPU box: tools/ab_sync_variants.sh run
set -e
cd "$(dirname "$0")/.."
PKG=marl_optimal_execution_b200
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
if [ "$1" = build ]; then
  mkdir -p build/ab
  nvcc $FL -DABX_FENCE_ONLY -o build/ab/fence_only.so $PKG/csrc/abx_sim.cu $PKG/csrc/abx_qnet.cu
  exit 0
fi
mkdir -p gpurun_out
for v in default strict fence_only default strict fence_only; do
  case $v in default) L=$PKG/libabides_b200.so;; strict) L=$PKG/libabides_b200_strict.so;; *) L=build/ab/fence_only.so;; esac
  echo "== $v" | tee -a gpurun_out/ab_sync.log
  ABX_LIB_PATH=$PWD/$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-rmsc03 --no-rmsc01 --no-ddqn --env-steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('lob msgs/s %.4g  env steps/s %.4g' % (d['value'], d['env']['value']))" | tee -a gpurun_out/ab_sync.log
done
