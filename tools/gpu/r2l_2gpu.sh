#!/bin/sh
# round 2, call L (2 GPUs): communication stream at the highest priority, one NCCL group per slice; lane priority A/B;
# the GPU test-suite once more (interpolation fast path, clean-failure test)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r2l_pytest.txt 2>&1
tail -3 gpurun_out/r2l_pytest.txt
show() { python -c "
import json; d=json.loads(open('gpurun_out/$1.json').read().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), d['config'].get('gathered_equals_local_recompute'), {n: round(v['ms_per_step'],1) for n,v in d['kernels'].items() if v['ms_per_step'] > 50})" || tail -5 gpurun_out/$1.err; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2l_n1.json 2> gpurun_out/r2l_n1.err; show r2l_n1
WB_LANE_PRIO=1 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2l_n1_prio.json 2> gpurun_out/r2l_n1_prio.err; show r2l_n1_prio
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2l_n1_nolanes.json 2> gpurun_out/r2l_n1_nolanes.err; show r2l_n1_nolanes
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2l_n2.json 2> gpurun_out/r2l_n2.err; show r2l_n2
