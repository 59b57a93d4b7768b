#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_h.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_gpu_tests_h.log
timeout 1200 python tools/train_ddqn.py --envs-per-gpu 2368 --episodes 9 --out gpurun_out/r02_ddqn_learning_curve.json > gpurun_out/r02_train.log 2>&1; echo "train rc=$?"; tail -12 gpurun_out/r02_train.log
