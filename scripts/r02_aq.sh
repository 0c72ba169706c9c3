#!/bin/bash
# last verification of round 2: full GPU suite and the default bench line with the final build
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02aq_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02aq_pytest_gpu.log
python bench.py > gpurun_out/r02s_bench_final_1gpu.json 2> gpurun_out/r02aq_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r02s_bench_final_1gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['verify']['ok'], d['cpu_baseline']['value'], d['roofline']['frac'], d['clocks'])"
