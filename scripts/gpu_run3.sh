#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roughconductor" > gpurun_out/pytest_rc.log 2>&1; echo "rc $?" >> gpurun_out/pytest_rc.log
tail -60 gpurun_out/pytest_rc.log
bash scripts/bench_short.sh "--debug shade_tune=0" "--debug shade_tune=2" "--debug shade_tune=3" > gpurun_out/shade_tune.log 2>&1
cat gpurun_out/shade_tune.log
