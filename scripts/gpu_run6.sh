#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_module.py tests/test_examples_flow.py -m gpu -q > gpurun_out/pytest_host.log 2>&1; echo "rc $?" >> gpurun_out/pytest_host.log
tail -40 gpurun_out/pytest_host.log
timeout 600 python scripts/bench_cfg5.py 1 > gpurun_out/cfg5.log 2>&1; cat gpurun_out/cfg5.log
