#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_module.py -m gpu -q -k "inverse" > gpurun_out/pytest_inverse.log 2>&1; echo "rc $?" >> gpurun_out/pytest_inverse.log
tail -30 gpurun_out/pytest_inverse.log
python scripts/gpu_configure_time.py > gpurun_out/configure_time.log 2>&1; cat gpurun_out/configure_time.log
