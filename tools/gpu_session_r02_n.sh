#!/bin/bash
# round 2, session n: the remaining execution baselines (VWAP schedule, Passive, Aggressive) on the GPU, then the whole GPU suite
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "vwap or passive or aggressive" > gpurun_out/r02_gpu_tests_n_new.log 2>&1; echo "new pytest rc=$?"; tail -6 gpurun_out/r02_gpu_tests_n_new.log
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_n.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gpu_tests_n.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_n.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_n.log
