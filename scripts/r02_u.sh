#!/bin/bash
mkdir -p gpurun_out
for a in "--batch 8388608 --debug pipeline=0" "--batch 8388608 --debug pipeline=2" "--batch 4194304 --debug pipeline=2" "--debug pipeline=0" "--debug pipeline=2"; do timeout 300 bash scripts/bench_short.sh "--no-verify $a"; done 2>&1 | tee gpurun_out/r02u_pipeline_ab.log
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02u_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02u_pytest_gpu.log
