#!/bin/sh
# round 2, final call: the whole -m gpu suite, the driver-shaped bench lines (ours, reference arm, configs 2 and 4),
# launch list and full ncu captures of the heaviest kernels at the final state
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2p_pytest.txt 2>&1
tail -3 gpurun_out/r2p_pytest.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'coded', d['e2e'].get('coded',{}).get('value'))
print('roofline', d['roofline']); print('parity', json.dumps(d['parity'])); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_reference_arm.json 2> gpurun_out/r2p_bench_reference_arm.err
cat gpurun_out/r2p_bench_reference_arm.json | cut -c1-900
python bench.py --config 2 --steps 2 --warmup 3 --no-coded > gpurun_out/r2p_bench_config2.json 2> gpurun_out/r2p_bench_config2.err
python bench.py --config 4 --steps 2 --warmup 3 > gpurun_out/r2p_bench_config4.json 2> gpurun_out/r2p_bench_config4.err
python -c "
import json
for c in ('config2','config4'):
    d=json.loads(open('gpurun_out/r2p_bench_%s.json'%c).read().splitlines()[-1]); print(c, round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'parity ok', d['parity']['device_resident']['within_1e-6'], d['parity']['e2e_host_arrays']['within_1e-6'])"
KR='regex:^(rng_fill|scan_counts|ct_|d4c_|harvest_|band_|decimate_pass|nyquist|fir_plain|dio_|stonemask)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-lanes > gpurun_out/r2p_ncu_bench.log 2>&1
for k in band_fir_events_kernel d4c_body_kernel harvest_refine_chain_kernel ct_frame_kernel band_interp_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -c 1 -f -o gpurun_out/r2p_$k python bench.py --utts 32 --steps 1 --warmup 0 --no-e2e --no-cpu --no-lanes > gpurun_out/r2p_ncu_$k.log 2>&1
done
ls gpurun_out | grep r2p
