#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "renderC or cfg1 or albedo or vertex_gradients_interior or other_fixture or shards" > gpurun_out/pytest_simple.log 2>&1; echo "rc $?" >> gpurun_out/pytest_simple.log
tail -5 gpurun_out/pytest_simple.log
bash scripts/bench_short.sh "--debug shade_simple=0" "--debug shade_simple=1" "--debug shade_simple=1 --debug shade_tune=2" "--debug shade_simple=1 --debug shade_tune=4"  "--debug shade_simple=1 --debug shade_tune=5" > gpurun_out/shade_simple.log 2>&1; cat gpurun_out/shade_simple.log
