#!/bin/bash
# (historical: sort_mode=6, occlusion rays sorted apart, was removed after this run)
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 bash scripts/bench_short.sh "--no-verify --debug sort_mode=5" "--no-verify --debug sort_mode=6" "--no-verify --debug sort_mode=5" "--no-verify --debug sort_mode=6" 2>&1 | tee gpurun_out/r02ac_sort_ray_type.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trace or headline" 2>&1 | tail -3
