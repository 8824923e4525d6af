#!/bin/sh
# round 2, call D: candidate-list selection in D4C, warp-per-frame remove kernel, world_b200_analyze_batch (two lanes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2d_pytest.txt 2>&1
tail -3 gpurun_out/r2d_pytest.txt
show() { python -c "
import json,sys; d=json.loads(open('$1').read().splitlines()[-1]); k=d['kernels']
print('$2', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'sum_kernels', round(sum(v['ms_per_step'] for v in k.values()),1), {n: round(v['ms_per_step'],1) for n,v in k.items() if v['ms_per_step'] > 1})"; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2d_nolanes.json 2> gpurun_out/r2d_nolanes.err; show gpurun_out/r2d_nolanes.json nolanes
for s in 2 4 8; do
  WB_LANE_SLICES=$s python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2d_lanes$s.json 2> gpurun_out/r2d_lanes$s.err; show gpurun_out/r2d_lanes$s.json lanes$s
done
WB_CT_THREADS=64 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2d_ct64.json 2> gpurun_out/r2d_ct64.err; show gpurun_out/r2d_ct64.json ct64
python bench.py --config 2 --steps 2 --warmup 2 --no-coded > gpurun_out/r2d_config2.json 2> gpurun_out/r2d_config2.err; show gpurun_out/r2d_config2.json config2
python bench.py --config 4 --steps 2 --warmup 2 > gpurun_out/r2d_config4.json 2> gpurun_out/r2d_config4.err; show gpurun_out/r2d_config4.json config4
python -c "
import json
for c in ('config2','config4'):
    d=json.loads(open('gpurun_out/r2d_%s.json'%c).read().splitlines()[-1]); print(c, 'e2e', d.get('e2e',{}).get('value'), 'parity', json.dumps(d.get('parity')), 'cpu', d.get('cpu_baseline'))"
