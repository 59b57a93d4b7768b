#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_d.log
timeout 1500 tools/ab_opt_variants.sh run
