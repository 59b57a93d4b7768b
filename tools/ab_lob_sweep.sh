#!/bin/bash
# LOB msgs/s of library variants (base = libabides_b200.so, else variant_<name>.so) against environments per GPU: low counts expose per-warp latency.
for v in "$@"; do
  lib=$PWD/marl_optimal_execution_b200/variant_$v.so; [ "$v" = base ] && lib=$PWD/marl_optimal_execution_b200/libabides_b200.so
  for n in 592 2072 2368 16384; do
    ABX_LIB_PATH=$lib python bench.py --steps 6 --warmup 3 --envs-per-gpu $n --no-cpu-baseline --no-env --no-ddqn 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', $n, '%.4g msgs/s' % d['value'], '%.2f ms/step' % d['ms_per_step'], 'err', d['config']['error_envs'])"
  done
done
