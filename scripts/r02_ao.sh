#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
timeout 600 bash scripts/bench_short.sh "--no-verify" "--no-verify --debug shade_tune=7" "--no-verify --debug shade_tune=8" 2>&1 | tee gpurun_out/r02ao_resolve_occupancy.log
