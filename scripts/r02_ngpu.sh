#!/bin/bash
# usage: r02_ngpu.sh N [extra bench args...] — bench on N GPUs of one box with the defaults (pixel tiles of 8 rows, NCCL collectives enqueued by the library)
N=$1; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_${N}gpu_gpus.txt
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -1 || exit 1
run() {   # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 "$@" 2>gpurun_out/r02_${N}gpu_err.log | tail -1 > gpurun_out/r02_bench_${N}gpu${tag}.json
  python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu${tag}.json'))
print('$N GPUs $tag $*', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'C', round(d['config']['ms_renderC'],1), 'D', round(d['config']['ms_renderD_vjp'],1), 'coll/step', d['config']['collectives_per_step'], 'verify', d['verify'].get('ok'))"
}
run ""
for a in "$@"; do run "_$(echo "$a" | tr -d ' =-')" $a; done
tail -2 gpurun_out/r02_${N}gpu_err.log
