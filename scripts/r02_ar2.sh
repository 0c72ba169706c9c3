#!/bin/bash
mkdir -p gpurun_out
timeout 100 bash scripts/bench_short.sh "--no-verify --debug bvh_builder=1 --debug lbvh_leaf=1" 2>&1 | tee gpurun_out/r02ar_lbvh_leaf1.log
