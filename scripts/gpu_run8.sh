#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu_variant_check.py > gpurun_out/variant_check.log 2>&1; tail -8 gpurun_out/variant_check.log
bash scripts/bench_short.sh "--debug trace_ld256=0" "--debug trace_ld256=1" "--debug trace_sstack=12" "--debug trace_sstack=16" "--debug trace_sstack=24" > gpurun_out/trace_variants.log 2>&1
cat gpurun_out/trace_variants.log
