#!/bin/bash
# 4 GPUs: bench with the defaults (pixel tiles of 8 rows) — NCCL path at more than two ranks, scaling point
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02r_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 5 --warmup 3 2>gpurun_out/r02r_err.log | tail -1 > gpurun_out/r02r_bench_4gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02r_bench_4gpu.json'))
print('4 GPUs', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'C', round(d['config']['ms_renderC'],1), 'D', round(d['config']['ms_renderD_vjp'],1), 'coll/step', d['config']['collectives_per_step'], 'verify', d['verify'].get('ok'), d['verify'].get('sharded_vs_unsharded'))"
tail -3 gpurun_out/r02r_err.log
