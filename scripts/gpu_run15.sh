#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "rc $?" >> gpurun_out/pytest_gpu_full.log
tail -25 gpurun_out/pytest_gpu_full.log
bash scripts/bench_short.sh "" > gpurun_out/bench_short.log 2>&1; cat gpurun_out/bench_short.log
