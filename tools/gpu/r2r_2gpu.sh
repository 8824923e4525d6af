#!/bin/sh
# round 2, call R (2 GPUs): peer-to-peer push of finished slices over CUDA IPC mappings (copy engines) vs NCCL broadcasts
mkdir -p gpurun_out
show() { python -c "
import json; d=json.loads(open('gpurun_out/$1.json').read().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config'].get('gathered_equals_local_recompute'))" || tail -15 gpurun_out/$1.err; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2r_n2_p2p.json 2> gpurun_out/r2r_n2_p2p.err; show r2r_n2_p2p
WB_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2r_n2_nccl.json 2> gpurun_out/r2r_n2_nccl.err; show r2r_n2_nccl
timeout 200 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2r_n1.json 2> gpurun_out/r2r_n1.err; show r2r_n1
