#!/bin/bash
# end-of-round evidence: sensor tests, smoke, bench (with the CPU baseline), cfg3 / cfg5 timings, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "sensor" > gpurun_out/pytest_sensor_final.log 2>&1; echo "rc $?" >> gpurun_out/pytest_sensor_final.log; tail -4 gpurun_out/pytest_sensor_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log
timeout 600 python scripts/bench_cfg3.py 5 > gpurun_out/cfg3.log 2>&1; grep -v "^{" gpurun_out/cfg3.log
timeout 600 python scripts/bench_cfg5.py 1 > gpurun_out/cfg5.log 2>&1; grep -v "^{" gpurun_out/cfg5.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_perm -s 20 -c 3 -o gpurun_out/prof_k_trace_perm_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 12 -c 1 -o gpurun_out/prof_k_resolve_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r.log 2>&1
ls -la gpurun_out | tail -8
