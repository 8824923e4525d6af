#!/bin/sh
# GPU experiment for DESIGN.md 9 item 2 (run on the B200 box from the repo root):
#   1. parity of the chain refinement kernel against the compiled reference on the GPU,
#   2. bench with the default kernel and with WB_REFINE_CHAIN=1 (device-resident part only).
# Usage: gpurun --timeout 500 -- 'sh tests/fuzz/exp_refine_chain.sh > gpurun_out/exp_refine_chain.txt 2>&1'
set -x
python - <<'PY'
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
for c in 0 1; do
  if [ $c = 1 ]; then export WB_REFINE_CHAIN=1; else unset WB_REFINE_CHAIN; fi
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); k=d['kernels']
print('chain=$c value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), {n: round(v['ms_per_step'],1) for n,v in k.items() if 'refine' in n or 'sweep' in n})"
done
