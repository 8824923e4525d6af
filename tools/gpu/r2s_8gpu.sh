#!/bin/sh
# round 2, call S (8 GPUs): weak scaling with the peer-to-peer push; BASELINE config 5 as named
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python -c "
import json; d=json.loads(open('gpurun_out/$name.json').read().splitlines()[-1])
print('$name', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['workload'], d['config']['gathered_equals_local_recompute'])" || tail -8 gpurun_out/$name.err
}
PORT=29571; run r2s_n8
PORT=29572; run r2s_n8_config5 --config 5
