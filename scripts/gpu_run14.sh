#!/bin/bash
mkdir -p gpurun_out
bash scripts/bench_short.sh "--batch 33554432" "--batch 67108864" "--batch 8388608" > gpurun_out/batch_ab.log 2>&1; cat gpurun_out/batch_ab.log
