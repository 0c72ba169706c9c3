#!/bin/bash
# round 2, two GPUs: NCCL render test + bench with pixel tiles (blocks / interleaved) and sample shards
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_ref_golden.py -q -m gpu -k "two_ranks or boundary_segment" > gpurun_out/r02h_pytest.log 2>&1
tail -15 gpurun_out/r02h_pytest.log
for a in "--shard pixels" "--shard pixels --tile-rows 8" "--shard samples"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $a 2>gpurun_out/r02h_err.log | tail -1 > gpurun_out/r02h_tmp.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02h_tmp.json'))
print('[$a]', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'C', round(d['config']['ms_renderC'],1), 'D', round(d['config']['ms_renderD_vjp'],1), 'coll/step', d['config']['collectives_per_step'], 'verify', d['verify'].get('ok'), d['verify'].get('sharded_vs_unsharded'))"
  cat gpurun_out/r02h_tmp.json >> gpurun_out/r02h_bench_2gpu.jsonl
  tail -3 gpurun_out/r02h_err.log
done 2>&1 | tee gpurun_out/r02h_ab.log
