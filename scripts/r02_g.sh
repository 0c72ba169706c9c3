#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02g_pytest_gpu.log 2>&1
tail -25 gpurun_out/r02g_pytest_gpu.log
