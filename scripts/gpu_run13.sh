#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "renderC or cfg1 or albedo or vertex_gradients or shards or forward_mode" > gpurun_out/pytest_ev.log 2>&1; echo "rc $?" >> gpurun_out/pytest_ev.log
tail -5 gpurun_out/pytest_ev.log
bash scripts/bench_short.sh "--debug shade_tune=6" "" "--debug shade_tune=2" > gpurun_out/shade_ev.log 2>&1; cat gpurun_out/shade_ev.log
