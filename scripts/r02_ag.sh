#!/bin/bash
# (historical: rec_by_shade was a run-time A/B switch; k_shade is now the only writer of vertex records)
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 bash scripts/bench_short.sh "--no-verify --debug rec_by_shade=0" "--no-verify --debug rec_by_shade=1" "--no-verify --debug rec_by_shade=0" "--no-verify --debug rec_by_shade=1" 2>&1 | tee gpurun_out/r02ag_rec_by_shade.log
python bench.py --steps 2 --no-cpu-baseline --debug rec_by_shade=1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['verify']['ok'])"
