#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trace or renderC or headline or shards" 2>&1 | tail -3
timeout 600 bash scripts/bench_short.sh "--no-verify" "--no-verify --batch 8388608" 2>&1 | tee gpurun_out/r02aa_scatter.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sort -s 30 -c 12 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>/dev/null | grep -E "k_sort" | awk -F'","' '{print $5, $(NF)}' | tee -a gpurun_out/r02aa_scatter.log
