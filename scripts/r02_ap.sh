#!/bin/bash
# device LBVH first build: parity + configure time + traversal rate on its tree
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "lbvh or refit" 2>&1 | grep -E "passed|failed|configure with|Error|assert" | tee gpurun_out/r02ap_lbvh.log
python bench.py --steps 2 --no-cpu-baseline --debug bvh_builder=1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bvh_builder=1', round(d['value'],1), 'Grays/s', round(d['roofline']['Grays_per_s'],2), 'verify', d['verify']['ok'])" | tee -a gpurun_out/r02ap_lbvh.log
