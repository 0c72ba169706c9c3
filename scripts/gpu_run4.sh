#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
bash scripts/bench_short.sh "" > gpurun_out/bench_short.log 2>&1
cat gpurun_out/bench_short.log
