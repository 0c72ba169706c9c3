#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config cfg3 --steps 3 --no-cpu-baseline > gpurun_out/r02o_bench_cfg3.json 2> gpurun_out/r02o_cfg3.err; tail -c 1500 gpurun_out/r02o_bench_cfg3.json; tail -3 gpurun_out/r02o_cfg3.err
timeout 1500 python bench.py --config cfg5 --steps 1 --no-cpu-baseline > gpurun_out/r02o_bench_cfg5.json 2> gpurun_out/r02o_cfg5.err; tail -c 1500 gpurun_out/r02o_bench_cfg5.json; tail -3 gpurun_out/r02o_cfg5.err
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "extension_is_loaded" 2>&1 | tail -2
