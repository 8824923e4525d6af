#!/bin/sh
# round 2, call C: new FFT kernels (CheapTrick / D4C on the self-sorting padded FFT, D4C body in-place + slow list),
# chain refinement as default, bench.py with parity / configs; ncu launch list + full captures of the four heaviest kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2c_pytest.txt 2>&1
tail -3 gpurun_out/r2c_pytest.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)
print({n: round(v['ms_per_step'],1) for n,v in d['kernels'].items()})
print('parity', json.dumps(d.get('parity')))
print('cpu', d.get('cpu_baseline'))
PY
WB_D4C_FAT=1 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2c_bench_fat.json 2> gpurun_out/r2c_bench_fat.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_fat.json').read().splitlines()[-1]); k=d['kernels']
print('D4C_FAT ms/step', round(d['ms_per_step'],1), {n: round(v['ms_per_step'],1) for n,v in k.items() if 'd4c' in n})"
KR='regex:^(rng_fill|scan_counts|ct_|d4c_|harvest_|band_sweep|decimate_pass|nyquist|fir_plain|dio_|stonemask)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2c_ncu_bench.log 2>&1
for k in d4c_body_kernel ct_frame_kernel harvest_refine_chain_kernel band_sweep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -c 1 -f -o gpurun_out/r2c_$k python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2c_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
