#!/bin/bash
for rep in 1 2; do for spec in base:9472 sm18:9472 sm18:10656; do v=${spec%%:*}; n=${spec#*:}
  echo -n "$v envs=$n: "
  ABX_LIB_PATH=$PWD/build/ab/opt_$v.so python bench.py --steps 3 --warmup 3 --envs-per-gpu 2368 --no-cpu-baseline --no-rmsc03 --no-rmsc01 --no-ddqn --no-whole-day --no-marketreplay --env-envs-per-gpu $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('env %.4g steps/s err %d' % (d['env']['value'], d['env']['error_envs']))"
done; done
