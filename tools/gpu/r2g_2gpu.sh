#!/bin/sh
# round 2, call G (2 GPUs): the C-ABI NCCL path -- slice-wise broadcast under the compute (analyze_batch_allgather)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2g_n2.json 2> gpurun_out/r2g_n2.err
tail -3 gpurun_out/r2g_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_n2.json').read().splitlines()[-1])
print('N=2 ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['multi_gpu'], d['config']['gathered_equals_local_recompute'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2g_n2_nolanes.json 2> gpurun_out/r2g_n2_nolanes.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_n2_nolanes.json').read().splitlines()[-1])
print('N=2 no-lanes (gather after compute) ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config']['gathered_equals_local_recompute'])"
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2g_n1.json 2> gpurun_out/r2g_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_n1.json').read().splitlines()[-1])
print('N=1 ms/step', round(d['ms_per_step'],1), 'value', round(d['value']))"
