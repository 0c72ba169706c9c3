#!/bin/bash
# linearised reflectance adjoint: parity tests, then A/B on the headline configuration
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_host_module.py tests/test_examples_verbatim.py -q -m gpu -x -k "albedo or bitmap or headline or path1 or launches or multi_view or adam or shards or together" 2>&1 | tail -8
timeout 600 bash scripts/bench_short.sh "--debug adjoint_lin=0" "--debug adjoint_lin=1" 2>&1 | tee gpurun_out/r02x_adjoint_lin_ab.log
python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['verify'])"
