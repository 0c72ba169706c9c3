#!/bin/bash
# sorted-copy traversal against the permutation form; L2 fetch granularity hint
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sorted_copy or albedo or headline" 2>&1 | tail -3
timeout 900 bash scripts/bench_short.sh "--no-verify --debug sorted_copy=0" "--no-verify --debug sorted_copy=1" "--no-verify --debug sorted_copy=0" "--no-verify --debug sorted_copy=1" "--no-verify --debug l2_fetch=32" "--no-verify --debug l2_fetch=128" 2>&1 | tee gpurun_out/r02ae_sorted_copy.log
for a in "sorted_copy=0" "sorted_copy=1" "l2_fetch=32"; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_trace_stream|k_sort_scatter|k_resolve' -s 4 -c 6 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify --debug $a 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' -v a=$a '{print a, substr($5,1,28), $(NF-2), $(NF)}' | tee -a gpurun_out/r02ae_sorted_copy.log
done
