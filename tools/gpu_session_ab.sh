#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 tools/ab_opt_variants.sh run
