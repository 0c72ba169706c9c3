#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02v_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02v_pytest_gpu.log
