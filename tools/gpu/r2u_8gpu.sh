#!/bin/sh
# round 2, call U (8 GPUs): ten tapered slices in the gather path
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2u_n8.json 2> gpurun_out/r2u_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2u_n8.json').read().splitlines()[-1])
print('r2u_n8 ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['gathered_equals_local_recompute'])" || tail -8 gpurun_out/r2u_n8.err
