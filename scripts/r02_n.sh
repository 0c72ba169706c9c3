#!/bin/bash
mkdir -p gpurun_out
for a in "--debug rng_seed_table=0" "--debug rng_seed_table=1"; do timeout 300 bash scripts/bench_short.sh "--no-verify $a"; done 2>&1 | tee gpurun_out/r02n_rng_seed_ab.log
timeout 3000 python -m pytest tests -q -m gpu -x > gpurun_out/r02n_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02n_pytest_gpu.log
