#!/bin/bash
mkdir -p gpurun_out
for v in 0 3 4 5; do
  timeout 300 bash scripts/bench_short.sh "--debug trace_node_min=402 --debug sort_mode=$v"
done 2>&1 | tee gpurun_out/r02e_ab.log
