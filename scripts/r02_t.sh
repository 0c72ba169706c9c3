#!/bin/bash
# 8 GPUs: bench with the defaults (pixel tiles of 8 rows) — NCCL path at more than two ranks, scaling point
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02t_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02t_err.log | tail -1 > gpurun_out/r02t_bench_4gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02t_bench_4gpu.json'))
print('8 GPUs', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'C', round(d['config']['ms_renderC'],1), 'D', round(d['config']['ms_renderD_vjp'],1), 'coll/step', d['config']['collectives_per_step'], 'verify', d['verify'].get('ok'), d['verify'].get('sharded_vs_unsharded'))"
tail -3 gpurun_out/r02t_err.log
