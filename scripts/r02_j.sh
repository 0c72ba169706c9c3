#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02j_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02j_pytest_gpu.log
timeout 600 python scripts/gpu_configure_time.py 2>&1 | tee gpurun_out/r02j_configure_time.log
for a in "" "--batch 8388608" "--batch 4194304"; do timeout 300 bash scripts/bench_short.sh "--no-verify $a"; done 2>&1 | tee gpurun_out/r02j_batch_ab.log
