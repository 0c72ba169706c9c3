#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_module.py -m gpu -q -k "envmap or roughconductor" > gpurun_out/pytest_env.log 2>&1; echo "rc $?" >> gpurun_out/pytest_env.log
tail -30 gpurun_out/pytest_env.log
timeout 600 python scripts/bench_cfg5.py 1 > gpurun_out/cfg5.log 2>&1; cat gpurun_out/cfg5.log
