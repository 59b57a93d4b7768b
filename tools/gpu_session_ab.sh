#!/bin/bash
# one A/B session: AB_VARIANTS="a b c" [AB_TESTS="tests/test_gpu_philox_oracle.py ..." AB_TEST_LIB=name]
set -u
mkdir -p gpurun_out
if [ -n "${AB_TESTS:-}" ]; then
  ABX_LIB_PATH=$PWD/build/ab/opt_${AB_TEST_LIB}.so timeout -s KILL 900 python -m pytest $AB_TESTS -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
fi
timeout 1500 tools/ab_opt_variants.sh run $AB_VARIANTS
