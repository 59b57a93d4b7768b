#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_c.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_c.log
timeout 1500 tools/ab_opt_variants.sh run
ABX_LIB_PATH=$PWD/build/ab/opt_vde.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_philox_oracle.py tests/test_gpu_book.py -q -k "z100 or z1000 or sparse_zi or book or tape" > gpurun_out/r02_gpu_tests_optvde.log 2>&1; echo "optvde pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests_optvde.log
