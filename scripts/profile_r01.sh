#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list of a short bench run + one full capture of the dominant kernel
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 240 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_perm -s 20 -c 3 -o gpurun_out/prof_k_trace_perm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
