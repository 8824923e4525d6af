"""Prototype for DESIGN.md section 9 item 2 (CPU only, numpy): Harvest's band-pass filter
    taps[k] = nuttall(2h+1)[k] * cos(w (k - h)),   h = round(2 afs / f),   w = 2 pi f / afs      (harvest.cpp:99-110)
is EXACTLY a sum of seven rectangular-window cosine filters, because the Nuttall window is four cosines
of k - h:  nuttall[k] = a0 + a1 cos(pi (k-h)/h) + a2 cos(2 pi (k-h)/h) + a3 cos(3 pi (k-h)/h).  A
rectangular-window cosine filter is a sliding DFT bin, i.e. a difference of two prefix sums of
x[m] exp(-j w_i m): O(1) per output sample instead of O(taps).  This script checks the identity and the
rounding behaviour (tile-local prefix sums, as a 2048-sample CTA tile would do) against the direct FIR and
against a long-double FIR, on synthetic speech at the decimated rate."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synth_batch  # noqa: E402

A = (0.355768, 0.487396, 0.144232, 0.012604)


def taps_of(f, afs, dtype=np.float64):
    h = int(np.floor(2.0 * afs / f + 0.5))
    k = np.arange(2 * h + 1, dtype=dtype)
    u = k / dtype(2 * h)
    pi = dtype(np.pi) if dtype is np.float64 else np.longdouble("3.14159265358979323846264338327950288")
    win = A[0] - A[1] * np.cos(2 * pi * u) + A[2] * np.cos(4 * pi * u) - A[3] * np.cos(6 * pi * u)
    return h, win * np.cos(2 * pi * dtype(f) * (k - h) / dtype(afs))


def direct(x, taps, h):
    # out[q] = sum_k taps[k] * x[q + h - k]  (zero outside); the reference's extra one-sample shift is irrelevant here
    return np.convolve(x, taps)[h:h + len(x)]


def sliding(x, f, afs, h, tile=2048):
    n = len(x)
    xp = np.concatenate([np.zeros(h), x, np.zeros(h + tile)])
    out = np.zeros(n)
    w0 = 2 * np.pi * f / afs
    comps = [(A[0], w0)] + [(A[i] / 2, w0 + s * i * np.pi / h) for i in (1, 2, 3) for s in (1, -1)]
    for q0 in range(0, n, tile):
        seg = xp[q0:q0 + tile + 2 * h]                       # samples q0-h .. q0+tile+h-1
        m = np.arange(len(seg))
        acc = np.zeros(tile)
        for c, w in comps:
            z = seg * np.exp(-1j * w * m)
            P = np.concatenate([[0.0], np.cumsum(z)])        # P[i] = sum_{m < i}
            q = np.arange(tile)
            D = P[q + 2 * h + 1] - P[q]                      # window [q, q+2h] of the segment
            acc += c * np.real(np.exp(1j * w * (q + h)) * D)
        out[q0:q0 + tile] = acc[:min(tile, n - q0)]
    return out


def crossings(y):
    i = np.nonzero((y[:-1] > 0) & (y[1:] <= 0))[0]
    return (i + 1) - y[i] / (y[i + 1] - y[i])


def main():
    afs = 8000.0
    x = synth_batch([1], 16000, 160000).numpy()[0][::2].copy()   # crude 8 kHz stand-in for the decimated signal
    x -= x.mean()
    print(f"{len(x)} samples at {afs:.0f} Hz")
    print(" f [Hz]  taps   |sliding-direct|/max|y|   |direct-exact|/max|y|   |sliding-exact|/max|y|   crossings   max crossing shift [samples]")
    for f in (40.0, 55.0, 80.0, 107.0, 160.0, 320.0):
        h, taps = taps_of(f, afs)
        yd = direct(x, taps, h)
        ys = sliding(x, f, afs, h)
        _, taps_l = taps_of(f, afs, np.longdouble)
        ye = np.convolve(x.astype(np.longdouble), taps_l)[h:h + len(x)]
        scale = np.abs(yd).max()
        cd, cs = crossings(yd), crossings(ys)
        shift = np.abs(cd - cs).max() if len(cd) == len(cs) else float("nan")
        print(f"{f:7.1f} {2 * h + 1:5d}   {np.abs(ys - yd).max() / scale:22.2e}   {float(np.abs(yd - ye).max()) / scale:21.2e}"
              f"   {float(np.abs(ys - ye).max()) / scale:21.2e}   {len(cd):9d}   {shift:12.2e}")


if __name__ == "__main__":
    main()
