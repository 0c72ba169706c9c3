#!/bin/bash
# smoke() on the box, then ncu --set full of the final shading-side kernels (renderD's first batch) for profiles/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify"
ncu --set full --clock-control none --import-source on -k regex:'k_resolve|k_shade|k_sort_scatter|k_adjoint' -s 100 -c 8 -o gpurun_out/r02am_shading_final $B > gpurun_out/r02am_ncu.log 2>&1
ls -la gpurun_out/r02am_* | tail -3
