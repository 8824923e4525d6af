#!/bin/sh
# round 2, final state: the whole -m gpu suite, the driver-shaped bench line, the ncu launch list of the same command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2y_pytest.txt 2>&1
tail -3 gpurun_out/r2y_pytest.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench_n1.json').read().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'coded', d['e2e'].get('coded',{}).get('value'))
print('roofline', d['roofline']); print('parity', json.dumps(d['parity'])[:900]); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d.get('gpu_launches'))
PY
KR='regex:^(rng_fill|scan_counts|ct_|d4c_|harvest_|band_|decimate_pass|nyquist|fir_plain|dio_|stonemask)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-lanes --cpu-utts 1 --parity-utts 0 > gpurun_out/r2y_ncu_bench.log 2>&1
ls -la gpurun_out | grep r2y
