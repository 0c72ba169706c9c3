#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "envmap or roughconductor" > gpurun_out/pytest_env.log 2>&1; echo "rc $?" >> gpurun_out/pytest_env.log
tail -60 gpurun_out/pytest_env.log
python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log
