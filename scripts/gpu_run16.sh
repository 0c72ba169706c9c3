#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_full2.log 2>&1; echo "rc $?" >> gpurun_out/pytest_gpu_full2.log
tail -6 gpurun_out/pytest_gpu_full2.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
bash scripts/bench_short.sh "" 2>&1 | tail -1
