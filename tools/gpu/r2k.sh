#!/bin/sh
# round 2, call K: leaner event detection in the FIR/event kernel; chain refinement at 5 CTAs/SM (A/B)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2k_pytest.txt 2>&1
tail -3 gpurun_out/r2k_pytest.txt
show() { python -c "
import json,sys; d=json.loads(open('$1').read().splitlines()[-1]); k=d['kernels']
print('$2', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'sum_kernels', round(sum(v['ms_per_step'] for v in k.values()),1), {n: round(v['ms_per_step'],1) for n,v in k.items() if v['ms_per_step'] > 1})"; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2k_split.json 2> gpurun_out/r2k_split.err; show gpurun_out/r2k_split.json split
WB_REFINE_OCC=5 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2k_occ5.json 2> gpurun_out/r2k_occ5.err; show gpurun_out/r2k_occ5.json refine_occ5
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2k_lanes.json 2> gpurun_out/r2k_lanes.err; show gpurun_out/r2k_lanes.json lanes
for k in band_fir_events_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -c 1 -f -o gpurun_out/r2k_$k python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2k_ncu_$k.log 2>&1
done
ls gpurun_out | grep r2k
