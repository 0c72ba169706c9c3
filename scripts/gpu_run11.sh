#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trace_bit or renderC or refit or cfg1 or vertex_gradients_interior or albedo" > gpurun_out/pytest_arena.log 2>&1; echo "rc $?" >> gpurun_out/pytest_arena.log
tail -8 gpurun_out/pytest_arena.log
bash scripts/bench_short.sh "--debug l2_persist=0" "--debug l2_persist=1" "--debug l2_persist=0" "--debug l2_persist=1" > gpurun_out/l2_persist.log 2>&1; cat gpurun_out/l2_persist.log
