#!/bin/sh
# round 2, call J (8 GPUs): weak scaling through the C-ABI NCCL path (slice-wise broadcast under the compute), then
# BASELINE.json configs[4] as named (8192 x 5 s sharded over 8 GPUs)
mkdir -p gpurun_out
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python -c "
import json; d=json.loads(open('gpurun_out/$name.json').read().splitlines()[-1])
print('$name', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['workload'], d['config']['gathered_equals_local_recompute'], {n: round(v['ms_per_step'],1) for n,v in d['kernels'].items() if v['ms_per_step'] > 20})" || tail -5 gpurun_out/$name.err
}
PORT=29521; run r2j_n8
PORT=29522; run r2j_n8_config5 --config 5
export WB_LANE_SLICES=4; PORT=29523; run r2j_n8_slices4
