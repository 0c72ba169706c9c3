#!/bin/bash
# launch list of a short bench run + ncu --set full of the traversal kernel (traffic) with the current build
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_perm -s 20 -c 3 -o gpurun_out/prof_k_trace_perm_c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_c2.log 2>&1
ls -la gpurun_out | tail -5
