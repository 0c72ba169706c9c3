#!/bin/bash
# round 2: full GPU test suite + default bench with the new traversal kernel, module e2e leg and verification
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 3000 gpurun_out/r02f_bench.json; tail -5 gpurun_out/r02f_bench.err
