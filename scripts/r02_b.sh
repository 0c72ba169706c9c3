#!/bin/bash
# round 2, second GPU call: prefetch / occupancy / refill-threshold variants of the streaming traversal kernel
mkdir -p gpurun_out
for v in 16 201 202 203 210 212 213 304 312 14 18; do
  timeout 300 bash scripts/bench_short.sh "--debug trace_kernel=3 --debug trace_node_min=$v"
done 2>&1 | tee gpurun_out/r02b_ab.log
