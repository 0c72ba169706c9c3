#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -q -m gpu -x -k "vertex or together or sorted_copy or reverse_mode or albedo or golden" 2>&1 | tail -3
for c in cfg3 cfg5; do python bench.py --config $c --no-cpu-baseline --steps 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$c', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1))"; done | tee gpurun_out/r02al_adjoint_shortcut.log
