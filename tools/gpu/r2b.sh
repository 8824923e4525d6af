#!/bin/sh
# round 2, call A: state of HEAD on a B200 + the two unmeasured items of DESIGN.md 9 (chain refinement, DIO path)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2b_pytest.txt 2>&1
tail -3 gpurun_out/r2b_pytest.txt
python - > gpurun_out/r2b_chain_parity.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from refworld import RefWorld, rel_err
from world_b200.api import World
from synth import synth_batch
ref = RefWorld(); w = World(device=0)
for fs, n, seeds in ((16000, 48000, [31, 32, 33]), (48000, 48000, [35])):
    x = synth_batch(seeds, fs, n, device="cuda:0")
    for chain in (0, 1):
        if chain: os.environ["WB_REFINE_CHAIN"] = "1"
        else: os.environ.pop("WB_REFINE_CHAIN", None)
        t, f0, fl = w.harvest(x, fs); w.synchronize()
        worst, flips = 0.0, 0
        for u in range(len(seeds)):
            tr, fr = ref.harvest(x[u].cpu().numpy(), fs)
            g = f0[u].cpu().numpy()[:fl[u]]
            flips += int(((g > 0) != (fr > 0)).sum())
            worst = max(worst, rel_err(g, fr).max() if not ((g > 0) != (fr > 0)).any() else float("inf"))
        print(f"fs {fs} chain {chain}: max rel err {worst:.2e}, V/UV flips {flips}", flush=True)
PY
cat gpurun_out/r2b_chain_parity.txt
for c in 0 1; do
  if [ $c = 1 ]; then export WB_REFINE_CHAIN=1; else unset WB_REFINE_CHAIN; fi
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_chain$c.json 2> gpurun_out/r2b_bench_chain$c.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/r2b_bench_chain$c.json').read().splitlines()[-1]); k=d['kernels']
print('chain=$c value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), {n: round(v['ms_per_step'],1) for n,v in k.items()})"
done
unset WB_REFINE_CHAIN
python bench.py --f0 dio --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_dio.json 2> gpurun_out/r2b_bench_dio.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r2b_bench_dio.json').read().splitlines()[-1]); k=d['kernels']
print('dio value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), {n: round(v['ms_per_step'],1) for n,v in k.items()})"
