#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "refit or configure or trace_bit" > gpurun_out/pytest_refit.log 2>&1; echo "rc $?" >> gpurun_out/pytest_refit.log
tail -30 gpurun_out/pytest_refit.log
