#!/bin/bash
# device LBVH: triangles per leaf
mkdir -p gpurun_out
timeout 200 bash scripts/bench_short.sh "--no-verify --debug bvh_builder=1 --debug lbvh_leaf=8" "--no-verify --debug bvh_builder=1 --debug lbvh_leaf=2" 2>&1 | tee gpurun_out/r02ar_lbvh_leaf.log
