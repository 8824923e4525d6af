#!/bin/sh
# round 2, call H: warp-specialised FIR/event kernel + binary-search interpolation; D4C / CheapTrick instruction cuts
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2h_pytest.txt 2>&1
tail -3 gpurun_out/r2h_pytest.txt
show() { python -c "
import json,sys; d=json.loads(open('$1').read().splitlines()[-1]); k=d['kernels']
print('$2', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'sum_kernels', round(sum(v['ms_per_step'] for v in k.values()),1), {n: round(v['ms_per_step'],1) for n,v in k.items() if v['ms_per_step'] > 1})"; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2h_split.json 2> gpurun_out/r2h_split.err; show gpurun_out/r2h_split.json split
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2h_lanes.json 2> gpurun_out/r2h_lanes.err; show gpurun_out/r2h_lanes.json lanes
python bench.py --config 4 --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2h_config4.json 2> gpurun_out/r2h_config4.err; show gpurun_out/r2h_config4.json config4
for k in band_fir_events_kernel band_interp_kernel d4c_body_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -c 1 -f -o gpurun_out/r2h_$k python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2h_ncu_$k.log 2>&1
done
ls gpurun_out | grep r2h
