"""Deterministic synthetic speech for tests and bench (same recipe in spirit):
per-utterance seed; base f0 ~ U(90, 250) Hz with 0.3-0.8 Hz vibrato (depth 25 %); every 1 s period is
0.75 s voiced (20 ms linear ramps) + 0.25 s unvoiced; voiced part = harmonic sum shaped by two
formant-like resonances; white noise 0.003*U(-1,1) on every sample (no digital silence).
Written with torch ops so the same code builds a [N, samples] float64 batch on the GPU (bench) or
on the CPU (tests); the array handed to the GPU path and to the CPU oracle is the same one."""
import math

import numpy as np
import torch


def synth_batch(seeds, fs, n_samples, device="cpu", zero_tail=0):
    seeds = list(seeds)
    n = len(seeds)
    base = np.empty(n); rate = np.empty(n)
    for i, s in enumerate(seeds):
        r = np.random.RandomState((1000003 * int(s) + 17) % (1 << 32))
        base[i] = r.uniform(90.0, 250.0)
        rate[i] = r.uniform(0.3, 0.8)
    dev = torch.device(device)
    t = torch.arange(n_samples, dtype=torch.float64, device=dev) / fs
    base_t = torch.tensor(base, dtype=torch.float64, device=dev)[:, None]
    rate_t = torch.tensor(rate, dtype=torch.float64, device=dev)[:, None]
    f0 = base_t * (1.0 + 0.25 * torch.sin(2 * math.pi * rate_t * t[None, :]))
    phase = torch.cumsum(2 * math.pi * f0 / fs, dim=1)
    tm = torch.remainder(t, 1.0)
    env = torch.clamp(torch.minimum(tm / 0.02, (0.75 - tm) / 0.02), 0.0, 1.0)[None, :]
    x = torch.zeros((n, n_samples), dtype=torch.float64, device=dev)
    kmax = int(min(40, math.floor(0.45 * fs / (base.max() * 1.25))))
    for k in range(1, kmax + 1):
        fk = k * f0
        g = 1.0 / (1.0 + ((fk - 700.0) / 300.0) ** 2) + 0.5 / (1.0 + ((fk - 1800.0) / 400.0) ** 2) + 0.05
        live = (fk < 0.45 * fs).to(torch.float64)
        x += live * 2.0 * g * torch.sin(k * phase) / k
    x *= 0.25 * env
    gen = torch.Generator(device=dev)
    gen.manual_seed(424242 + int(seeds[0]))
    x += 0.003 * (2.0 * torch.rand((n, n_samples), dtype=torch.float64, device=dev, generator=gen) - 1.0)
    if zero_tail:
        x[:, n_samples - zero_tail:] = 0.0
    return x.contiguous()
