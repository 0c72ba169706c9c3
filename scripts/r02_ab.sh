#!/bin/bash
# (historical: scatter_items was a run-time A/B switch)
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 bash scripts/bench_short.sh "--no-verify --debug scatter_items=8" "--no-verify --debug scatter_items=1604" "--no-verify --debug scatter_items=816" "--no-verify --debug scatter_items=808" "--no-verify --debug scatter_items=432" "--no-verify --debug scatter_items=416" 2>&1 | tee gpurun_out/r02ab_scatter_shapes.log
