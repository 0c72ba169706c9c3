#!/bin/bash
# ncu captures of the streaming traversal kernel: full 32 Mi-lane wavefront and the 8 Mi-lane wavefront one GPU of an 8-GPU job sees
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_trace_stream -s 20 -c 2 -o gpurun_out/r02l_stream_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r02l_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_stream -s 60 -c 2 -o gpurun_out/r02l_stream_8m python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify --batch 8388608 > gpurun_out/r02l_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r02l_ncu3.log 2>&1
timeout 600 python scripts/gpu_configure_time.py 2>&1 | tail -2
ls -la gpurun_out | tail -6
