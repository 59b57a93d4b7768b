#!/bin/bash
# A/B of step-kernel variants (marl_optimal_execution_b200/variant_<name>.so): ABIDESEnv episode + DDQN ticks of bench.py, LOB block skipped.
for v in "$@"; do
  lib=$PWD/marl_optimal_execution_b200/variant_$v.so; [ "$v" = base ] && lib=$PWD/marl_optimal_execution_b200/libabides_b200.so
  ABX_LIB_PATH=$lib python bench.py --steps 3 --warmup 3 --envs-per-gpu 2048 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'env %.4g steps/s' % d['env']['value'], 'ddqn %.4g ticks/s' % d['ddqn']['value'], 'train %.4g' % d['ddqn']['training']['value'])"
done
