#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sorted_copy or albedo or headline or launches or bitmap" 2>&1 | tail -3
timeout 600 bash scripts/bench_short.sh "--no-verify --debug sorted_copy=1" "--no-verify" 2>&1 | tee gpurun_out/r02ah_sorted_copy_after_records.log
