#!/bin/bash
# round 2, session o: the streamed Q-network kernel (NNModel_2)
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_qnet.py -m gpu -q -x > gpurun_out/r02_gpu_tests_o.log 2>&1; echo "qnet pytest rc=$?"; tail -15 gpurun_out/r02_gpu_tests_o.log
