#!/bin/bash
# ncu --set full of the shading-side kernels (renderD's first batch: k_primary, k_shade, k_resolve; the linearised adjoint; the sort passes)
# and the launch list of the step with the linearised adjoint
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify"
ncu --set full --clock-control none --import-source on -k regex:'k_resolve|k_shade|k_primary|k_adjoint' -s 44 -c 5 -o gpurun_out/r02y_shading $B > gpurun_out/r02y_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_resolve|k_shade|k_primary|k_adjoint' -s 88 -c 2 -o gpurun_out/r02y_adjoint_lin $B > gpurun_out/r02y_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sort_scatter|k_sort_hist|k_trace_compact' -s 9 -c 3 -o gpurun_out/r02y_sort $B > gpurun_out/r02y_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/r02y_ncu4.log 2>&1
ls -la gpurun_out | tail -8
