#!/bin/bash
# usage: bench_short.sh "<bench args>" ["<bench args>" ...]
for a in "$@"; do
python bench.py --steps 3 --no-cpu-baseline $a 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$a]', round(d['value'],1), 'C ms', round(d['config']['ms_renderC'],1), 'D+vjp ms', round(d['config']['ms_renderD_vjp'],1), 'Grays/s', round(d['roofline']['Grays_per_s'],2), 'frac', round(d['roofline']['frac'],4))"
done
