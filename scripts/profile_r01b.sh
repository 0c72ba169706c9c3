#!/bin/bash
# round-1 re-entry: GPU tests, short bench, ncu --set full of the shading-side kernels (k_resolve, k_shade, k_adjoint, k_primary)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1
for k in k_resolve k_adjoint k_shade k_sort_scatter; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench.log
