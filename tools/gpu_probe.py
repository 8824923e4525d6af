"""Ad-hoc per-stage timing on one GPU (development aid; bench.py is the contract)."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from world_b200.api import World
from synth import synth_batch


def true_f0(seeds, fs, n_frames, frame_period=5.0):
    out = np.zeros((len(seeds), n_frames))
    t = np.arange(n_frames) * frame_period / 1000.0
    for i, s in enumerate(seeds):
        r = np.random.RandomState((1000003 * int(s) + 17) % (1 << 32))
        base = r.uniform(90.0, 250.0); rate = r.uniform(0.3, 0.8)
        f = base * (1.0 + 0.25 * np.sin(2 * math.pi * rate * t))
        f[np.mod(t, 1.0) >= 0.75] = 0.0
        out[i] = f
    return out


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    fs = int(os.environ.get("FS", 16000)); n_utts = int(os.environ.get("N", 128)); sec = float(os.environ.get("SEC", 10))
    w = World(device=0)
    seeds = list(range(1, n_utts + 1))
    x = synth_batch(seeds, fs, int(fs * sec), device="cuda:0")
    L = w.frames(fs, x.shape[1])
    t = torch.arange(L, dtype=torch.float64, device="cuda:0")[None, :].repeat(n_utts, 1) * 5 / 1000.0
    f0 = torch.from_numpy(true_f0(seeds, fs, L)).cuda()
    opt = w.cheaptrick_option(fs)
    bins = opt.fft_size // 2 + 1
    sp = torch.empty((n_utts, L, bins), dtype=torch.float64, device="cuda:0")
    ap = torch.empty_like(sp)
    frames = n_utts * L
    res = {}
    res["stonemask"] = timeit(lambda: w.stonemask(x, fs, t, f0))
    res["cheaptrick"] = timeit(lambda: w.cheaptrick(x, fs, t, f0, opt, out=sp))
    res["d4c"] = timeit(lambda: w.d4c(x, fs, t, f0, opt.fft_size, out=ap))
    for name in ("dio", "harvest"):
        try:
            res[name] = timeit(lambda: getattr(w, name)(x, fs), n=2)
        except Exception as e:  # stage not built yet
            print(name, "unavailable:", str(e)[:80])
    w.synchronize()
    for k, v in res.items():
        print(f"{k:12s} {v:9.2f} ms  {frames / v * 1e3 / 1e6:8.2f} Mframes/s")
    print("frames", frames, "fs", fs, "n", n_utts)


if __name__ == "__main__":
    main()
