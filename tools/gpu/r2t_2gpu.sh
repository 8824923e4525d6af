#!/bin/sh
# round 2, call T (2 GPUs): tapered slices in the gather path
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2t_n2.json 2> gpurun_out/r2t_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2t_n2.json').read().splitlines()[-1])
print('r2t_n2 ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config'].get('gathered_equals_local_recompute'))" || tail -15 gpurun_out/r2t_n2.err
