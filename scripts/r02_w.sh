#!/bin/bash
# final round-2 single-GPU numbers: the default bench line, the reference arm, the other configurations
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
python bench.py > gpurun_out/r02s_bench_final_1gpu.json 2> gpurun_out/r02w_err.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_bench_reference_arm.json 2>> gpurun_out/r02w_err.log
for c in cfg3 cfg4 cfg5; do python bench.py --config $c --no-cpu-baseline > gpurun_out/r02w_bench_$c.json 2>> gpurun_out/r02w_err.log; done
timeout 300 bash scripts/bench_short.sh "--no-verify --debug shade_tune=0" "--no-verify --debug shade_tune=2" 2>&1 | tee gpurun_out/r02w_shade_tune_ab.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02s_bench_final_1gpu.json')+glob.glob('gpurun_out/r02w_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d.get('verify',{}).get('ok'), d.get('cpu_baseline',{}).get('value'), d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r02w_err.log
