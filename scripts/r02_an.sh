#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
timeout 600 bash scripts/bench_short.sh "--no-verify" "--no-verify" 2>&1 | tee gpurun_out/r02an_stg256.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trace or headline or albedo or vertex_gradients_all or sorted" 2>&1 | tail -2
