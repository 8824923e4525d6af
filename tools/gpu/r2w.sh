#!/bin/sh
# round 2: A/B of (1) quarter-size sub-chunks in the last outer chunk of analyze_host, (2) flattened window loads in
# band_interp_kernel, (3) out-of-line smoothing helpers in the D4C body -- variants as libworld_b200_<tag>.so
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_fuzz.py -m gpu -x -q -k "harvest or chunking or dense or fuzz or analyze" > gpurun_out/r2w_pytest.txt 2>&1
tail -2 gpurun_out/r2w_pytest.txt
B="python bench.py --steps 3 --warmup 2 --no-cpu --cpu-utts 1 --parity-utts 0"
$B --no-coded > gpurun_out/r2w_taper.json 2> gpurun_out/r2w_taper.err
WB_HOST_TAPER=0 $B --no-coded > gpurun_out/r2w_notaper.json 2> gpurun_out/r2w_notaper.err
$B --no-e2e --no-lanes > gpurun_out/r2w_flat.json 2> gpurun_out/r2w_flat.err
WORLD_B200_LIB=$PWD/world_b200/lib/libworld_b200_ab.so $B --no-e2e --no-lanes > gpurun_out/r2w_ab.json 2> gpurun_out/r2w_ab.err
WORLD_B200_LIB=$PWD/world_b200/lib/libworld_b200_ab2.so $B --no-e2e --no-lanes > gpurun_out/r2w_ab2.json 2> gpurun_out/r2w_ab2.err
python - <<'PY'
import json
for f in ('taper','notaper','flat','ab','ab2'):
    try:
        d=json.loads(open('gpurun_out/r2w_%s.json'%f).read().splitlines()[-1])
        k=d.get('kernels') or {}
        def ms(name):
            v=k.get(name)
            return v if not isinstance(v,dict) else v.get('ms_per_step', v)
        print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']) if d.get('e2e') and d['e2e'].get('value') else None,
              'parity', d['parity']['device_resident']['within_1e-6'] if d.get('parity') else None)
        if isinstance(k,dict):
            print('   ', {n: ms(n) for n in k if any(t in n for t in ('interp','d4c_body','fir_events','refine_chain','ct_'))})
        else:
            print('   ', str(k)[:600])
    except Exception as e:
        print(f,'failed',e)
PY
