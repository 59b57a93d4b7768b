#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reset.py tests/test_gpu_env.py tests/test_gpu_ddqn.py -m gpu -q > gpurun_out/r02_gpu_tests_e.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_e.log
timeout 1500 tools/ab_opt_variants.sh run
