#!/bin/sh
# round 2: A/B of load-prefetch variants (in-place FFT twiddles, window samples / draws in D4C and CheapTrick) and of
# the e2e sub-chunk divisor
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft or golden or chunking" > gpurun_out/r2x_pytest.txt 2>&1
tail -2 gpurun_out/r2x_pytest.txt
for t in pa pb pc; do
  WORLD_B200_LIB=$PWD/world_b200/lib/libworld_b200_$t.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft or golden_cheaptrick" > gpurun_out/r2x_pytest_$t.txt 2>&1
  tail -1 gpurun_out/r2x_pytest_$t.txt
done
B="python bench.py --steps 3 --warmup 2 --no-cpu --cpu-utts 1 --parity-utts 0"
$B --no-e2e --no-lanes > gpurun_out/r2x_base.json 2> gpurun_out/r2x_base.err
for t in pa pb pc; do
  WORLD_B200_LIB=$PWD/world_b200/lib/libworld_b200_$t.so $B --no-e2e --no-lanes > gpurun_out/r2x_$t.json 2> gpurun_out/r2x_$t.err
done
WB_HOST_TAPER_DIV=8 $B --no-coded > gpurun_out/r2x_div8.json 2> gpurun_out/r2x_div8.err
python - <<'PY'
import json
for f in ('base','pa','pb','pc','div8'):
    try:
        d=json.loads(open('gpurun_out/r2x_%s.json'%f).read().splitlines()[-1])
        k=d.get('kernels') or {}
        print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']) if d.get('e2e') and d['e2e'].get('value') else None)
        print('   ', {n: round(v['ms_per_step'],2) for n,v in k.items() if any(t in n for t in ('d4c_body_kernel','d4c_lovetrain','ct_frame','interp'))})
    except Exception as e:
        print(f,'failed',e)
PY
