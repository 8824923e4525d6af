#!/bin/sh
# round 2, call F (E again after the shared-memory size fix): split sweep (band_fir_events_kernel with TMA staging + band_interp_kernel) vs the streaming kernel;
# D4C power row sized for the longest window (config 4)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2f_pytest.txt 2>&1
tail -3 gpurun_out/r2f_pytest.txt
show() { python -c "
import json,sys; d=json.loads(open('$1').read().splitlines()[-1]); k=d['kernels']
print('$2', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'sum_kernels', round(sum(v['ms_per_step'] for v in k.values()),1), {n: round(v['ms_per_step'],1) for n,v in k.items() if v['ms_per_step'] > 1})"; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2f_split.json 2> gpurun_out/r2f_split.err; show gpurun_out/r2f_split.json split
WB_SWEEP_STREAMING=1 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2f_streaming.json 2> gpurun_out/r2f_streaming.err; show gpurun_out/r2f_streaming.json streaming
python bench.py --config 4 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2f_config4.json 2> gpurun_out/r2f_config4.err; show gpurun_out/r2f_config4.json config4
for k in band_fir_events_kernel band_interp_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -c 1 -f -o gpurun_out/r2f_$k python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2f_ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:^d4c_body_kernel\$ -c 1 -f -o gpurun_out/r2f_d4c_body_kernel python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2f_ncu_d4c.log 2>&1
ls gpurun_out | grep r2f
