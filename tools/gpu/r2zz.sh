#!/bin/sh
# round 2, final state: smoke() and the config-2 line (Dio + StoneMask + CheapTrick + D4C) with the last kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zz_smoke.txt 2>&1; tail -2 gpurun_out/r2zz_smoke.txt
python bench.py --config 2 --steps 2 --warmup 3 --no-coded --no-cpu --cpu-utts 2 --parity-utts 1 > gpurun_out/r2zz_bench_config2.json 2> gpurun_out/r2zz_bench_config2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2zz_bench_config2.json').read().splitlines()[-1]); print('config2', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'parity', d['parity']['device_resident']['within_1e-6'], d['parity']['e2e_host_arrays']['within_1e-6'])"
