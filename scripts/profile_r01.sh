#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list of a short bench run + one full capture of k_trace
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 3 -o gpurun_out/prof_k_trace python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 40 -c 2 -o gpurun_out/prof_k_resolve python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu3.log 2>&1
ls -la gpurun_out
