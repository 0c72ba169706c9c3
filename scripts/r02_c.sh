#!/bin/bash
# round 2: ncu --set full of the streaming kernel and of the compact one-ray-per-thread kernel on the bench workload
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_trace_stream -s 20 -c 2 -o gpurun_out/r02c_stream python bench.py --steps 1 --warmup 3 --no-cpu-baseline --debug trace_kernel=3 --debug trace_node_min=312 > gpurun_out/r02c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_compact -s 20 -c 1 -o gpurun_out/r02c_compact python bench.py --steps 1 --warmup 3 --no-cpu-baseline --debug trace_kernel=1 > gpurun_out/r02c_ncu2.log 2>&1
ls -la gpurun_out | tail -5
