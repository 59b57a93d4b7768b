#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_b.log
timeout 1200 tools/ab_opt_variants.sh run
ABX_LIB_PATH=$PWD/build/ab/opt_all.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_philox_oracle.py -q -k "z100 or z1000 or sparse_zi" > gpurun_out/r02_gpu_tests_optall.log 2>&1; echo "optall pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests_optall.log
