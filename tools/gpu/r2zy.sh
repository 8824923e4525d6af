#!/bin/sh
# config 2 (download-bound pipeline): e2e with and without the quarter-size sub-chunks of the last outer chunk
mkdir -p gpurun_out
B="timeout 30 python bench.py --config 2 --steps 3 --warmup 2 --no-coded --no-cpu --cpu-utts 1 --parity-utts 0"
$B > gpurun_out/r2zy_c2_taper.json 2> /dev/null
WB_HOST_TAPER=0 $B > gpurun_out/r2zy_c2_notaper.json 2> /dev/null
python -c "
import json
for f in ('taper','notaper'):
    try:
        d=json.loads(open('gpurun_out/r2zy_c2_%s.json'%f).read().splitlines()[-1]); print(f, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step_rank0'])
    except Exception as e: print(f,'failed',e)"
