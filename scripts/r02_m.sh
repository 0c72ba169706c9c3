#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02m_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02m_pytest_gpu.log
for a in "--debug sort_mode=5" "--debug sort_mode=6"; do timeout 300 bash scripts/bench_short.sh "--no-verify $a"; done 2>&1 | tee gpurun_out/r02m_sort_ab.log
timeout 600 python scripts/gpu_configure_time.py 2>&1 | tail -2 | tee gpurun_out/r02m_configure_time.log
