#!/bin/sh
# round 2, call M: band-pair FIR/event kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=3 -k "harvest or fuzz or analyze or benchmark or event or edge" > gpurun_out/r2m_pytest.txt 2>&1
tail -3 gpurun_out/r2m_pytest.txt
show() { python -c "
import json; d=json.loads(open('gpurun_out/$1.json').read().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), {n: round(v['ms_per_step'],1) for n,v in d['kernels'].items() if v['ms_per_step'] > 5})" || tail -5 gpurun_out/$1.err; }
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-lanes > gpurun_out/r2m_nolanes.json 2> gpurun_out/r2m_nolanes.err; show r2m_nolanes
python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2m_lanes.json 2> gpurun_out/r2m_lanes.err; show r2m_lanes
ncu --set full --clock-control none --import-source on -k regex:^band_fir_events_kernel\$ -c 1 -f -o gpurun_out/r2m_band_fir_events_kernel python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2m_ncu.log 2>&1
