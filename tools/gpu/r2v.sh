#!/bin/sh
# round 2: codec fused into the frame kernels -- -m gpu suite, bench line (coded leg fused), coded leg unfused for A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2v_pytest.txt 2>&1
tail -3 gpurun_out/r2v_pytest.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err
WB_CODEC_UNFUSED=1 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2v_bench_unfused.json 2> gpurun_out/r2v_bench_unfused.err
python - <<'PY'
import json
for f in ('n1','unfused'):
    try:
        d=json.loads(open('gpurun_out/r2v_bench_%s.json'%f).read().splitlines()[-1])
        print(f,'value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'coded', d['e2e'].get('coded'))
        print(' parity ok', d['parity']['device_resident']['within_1e-6'], d['parity']['e2e_host_arrays']['within_1e-6'], 'clocks', d['clocks'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -5 gpurun_out/r2v_bench_n1.err
