#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 600 bash scripts/bench_short.sh "--no-verify" "--no-verify --debug shade_tune=2" "--no-verify" "--no-verify --debug shade_tune=2" 2>&1 | tee gpurun_out/r02ai_conn_records.log
python bench.py --steps 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['verify']['ok'])" | tee -a gpurun_out/r02ai_conn_records.log
timeout 3000 python -m pytest tests -q -m gpu -x > gpurun_out/r02ai_pytest_gpu.log 2>&1
tail -6 gpurun_out/r02ai_pytest_gpu.log
