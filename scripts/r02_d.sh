#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trace or path5 or path1" > gpurun_out/r02d_pytest.log 2>&1
tail -3 gpurun_out/r02d_pytest.log
for v in 312 402 16; do
  timeout 300 bash scripts/bench_short.sh "--debug trace_kernel=3 --debug trace_node_min=$v"
done 2>&1 | tee gpurun_out/r02d_ab.log
