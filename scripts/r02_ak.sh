#!/bin/bash
# launch list of BASELINE configs[2] (vertex gradients, all terms)
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv --log-file gpurun_out/r02ak_cfg3_launches.csv python bench.py --config cfg3 --steps 1 --warmup 2 --no-cpu-baseline --no-verify > /dev/null 2>&1
wc -l gpurun_out/r02ak_cfg3_launches.csv
