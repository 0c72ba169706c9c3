#!/bin/bash
mkdir -p gpurun_out
for a in "" "--debug trace_chunk=128" "--debug trace_chunk=64" "--debug trace_blocks=6" "--debug trace_blocks=4" "--debug trace_node_min=1"; do timeout 300 bash scripts/bench_short.sh "--no-verify --batch 8388608 $a"; done 2>&1 | tee gpurun_out/r02q_small_wavefront_ab.log
for a in "--debug trace_chunk=128" "--debug trace_blocks=6"; do timeout 300 bash scripts/bench_short.sh "--no-verify $a"; done 2>&1 | tee -a gpurun_out/r02q_small_wavefront_ab.log
