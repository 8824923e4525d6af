#!/bin/sh
# round 2, call Q (8 GPUs): fewer NCCL channels = fewer SMs taken from the compute kernels by the transfer
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python -c "
import json; d=json.loads(open('gpurun_out/$name.json').read().splitlines()[-1])
print('$name', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['gathered_equals_local_recompute'])" || tail -5 gpurun_out/$name.err
}
export NCCL_MAX_NCHANNELS=8; PORT=29551; run r2q_n8_ch8
export NCCL_MAX_NCHANNELS=4; PORT=29552; run r2q_n8_ch4
