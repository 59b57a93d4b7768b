#!/bin/bash
# A/B of kernel variants (marl_optimal_execution_b200/variant_*.so built with -DABX_NO_* switches): short LOB bench each.
for v in "$@"; do
  ABX_LIB_PATH=$PWD/marl_optimal_execution_b200/variant_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-env --no-ddqn 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '%.4g msgs/s' % d['value'], 'err', d['config']['error_envs'])"
done
