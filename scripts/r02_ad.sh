#!/bin/bash
# (historical: the l2_fetch debug key — cudaLimitMaxL2FetchGranularity — existed only for this run; result in profiles/r02ae_*: no effect)
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2 || exit 1
timeout 900 bash scripts/bench_short.sh "--no-verify --debug l2_fetch=32" "--no-verify --debug l2_fetch=64" "--no-verify --debug l2_fetch=128" 2>&1 | tee gpurun_out/r02ad_l2_fetch.log
for g in 32 128; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_trace_stream -s 6 -c 2 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify --debug l2_fetch=$g 2>/dev/null | grep -E "k_trace" | awk -F'","' -v g=$g '{print "l2_fetch=" g, $(NF-2), $(NF)}' | tee -a gpurun_out/r02ad_l2_fetch.log
done
