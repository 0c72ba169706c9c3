#!/bin/bash
# round 2, first GPU call: parity of the new traversal kernels + A/B of the variants on the bench workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trace or hit or table or render_c or path" > gpurun_out/r02a_pytest_trace.log 2>&1
tail -3 gpurun_out/r02a_pytest_trace.log
for v in "trace_kernel=0" "trace_kernel=1" "trace_kernel=2" "trace_kernel=3" "trace_kernel=3 --debug trace_node_min=8" "trace_kernel=3 --debug trace_node_min=12" "trace_kernel=3 --debug trace_node_min=16" "trace_kernel=3 --debug trace_node_min=20" "trace_kernel=3 --debug trace_node_min=101" "trace_kernel=3 --debug trace_node_min=116"; do
  timeout 300 bash scripts/bench_short.sh "--debug $v"
done 2>&1 | tee gpurun_out/r02a_ab.log
