#!/bin/bash
# final single-GPU numbers of round 2: full GPU suite, default bench line, reference arm, the other configurations, launch list
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 3000 python -m pytest tests -q -m gpu > gpurun_out/r02af_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02af_pytest_gpu.log
python bench.py > gpurun_out/r02s_bench_final_1gpu.json 2> gpurun_out/r02af_err.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02af_bench_reference_arm.json 2>> gpurun_out/r02af_err.log
for c in cfg3 cfg4 cfg5; do python bench.py --config $c --no-cpu-baseline > gpurun_out/r02af_bench_$c.json 2>> gpurun_out/r02af_err.log; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r02af_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02s_bench_final_1gpu.json')+glob.glob('gpurun_out/r02af_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), 'e2e', round(d['e2e']['value'],2), (d.get('verify') or {}).get('ok'), (d.get('cpu_baseline') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02af_err.log
