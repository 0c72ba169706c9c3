#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu -x > gpurun_out/r02p_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02p_pytest_gpu.log
timeout 300 bash scripts/bench_short.sh "--no-verify" 2>&1 | tee gpurun_out/r02p_ab.log
timeout 900 python bench.py --config cfg3 --steps 3 --no-cpu-baseline > gpurun_out/r02p_bench_cfg3.json 2> gpurun_out/r02p_cfg3.err; python -c "
import json
d=json.loads(open('gpurun_out/r02p_bench_cfg3.json').read().strip().splitlines()[-1])
print('cfg3 value %.1f'%d['value'], 'e2e %.1f'%d['e2e']['value'], 'C %.1f D %.1f'%(d['config']['ms_renderC'], d['config']['ms_renderD_vjp']))" | tee -a gpurun_out/r02p_ab.log
