"""Prototype for DESIGN.md section 9 item 2 (CPU only, numpy): at a decimated rate of 8000 Hz one 1 ms frame
is exactly 8 samples, so Harvest's seven overlapped refinements of one base candidate (harvest.cpp:417-429,
589-617) use the same window and the same twiddles on 8-sample-shifted segments.  Compares the refined f0 /
score computed the reference's way (window rebuilt per frame) with the shared-template way."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synth_batch  # noqa: E402


def rnd(x):
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def fix_f0(main, diff, f, fs, nfft, H):
    num = den = score = 0.0
    for m in range(H):
        k = rnd(f * nfft / fs * (m + 1))
        pw = abs(main[m]) ** 2
        ni = main[m].real * diff[m].imag - main[m].imag * diff[m].real
        inst = 0.0 if pw == 0 else k * fs / nfft + ni / pw * fs / 2 / np.pi
        amp = np.sqrt(pw)
        num += amp * inst; den += amp * (m + 1.0); score += abs((inst / (m + 1.0) - f) / f)
    return num / (den + 1e-12), 1.0 / (score / H + 1e-12)


def main():
    fs = 8000.0
    y = synth_batch([3], 16000, 64000).numpy()[0][::2].copy()
    y -= y.mean()
    n = len(y)
    rng = np.random.default_rng(0)
    worst_f = worst_s = 0.0
    for trial in range(200):
        f = rng.uniform(72, 700)
        k0 = int(rng.integers(100, 3800))
        h = int(1.5 * fs / f + 1.0)
        nwin = 2 * h + 1
        T = (2.0 * h + 1.0) / fs
        nfft = 2 ** (2 + int(np.log(2.0 * h + 1.0) / 0.69314718055994529))
        H = min(int(fs / 2.0 / f), 6)
        bins = [rnd(f * nfft / fs * (m + 1)) for m in range(H)]
        j = np.arange(nwin)
        # shared template: window argument (j - h - 1) / fs, twiddles exp(-j 2 pi bin j / nfft)
        tau0 = (j - h - 1.0) / fs
        w0 = 0.42 + 0.5 * np.cos(2 * np.pi * tau0 / T) + 0.08 * np.cos(4 * np.pi * tau0 / T)
        dw0 = np.empty(nwin); dw0[0] = -w0[1] / 2; dw0[-1] = w0[-2] / 2; dw0[1:-1] = -(w0[2:] - w0[:-2]) / 2
        tw = np.exp(-2j * np.pi * np.outer(bins, j) / nfft)
        for g in range(-3, 4):
            k = k0 + g
            t = k * 1 / 1000.0
            basic = rnd((t + (-h) / fs) * fs + 0.001)
            assert basic == 8 * k - h
            tau = (basic + j - 1.0) / fs - t                       # the reference's expression, per frame
            w = 0.42 + 0.5 * np.cos(2 * np.pi * tau / T) + 0.08 * np.cos(4 * np.pi * tau / T)
            dw = np.empty(nwin); dw[0] = -w[1] / 2; dw[-1] = w[-2] / 2; dw[1:-1] = -(w[2:] - w[:-2]) / 2
            seg = y[np.clip(basic + j - 1, 0, n - 1)]
            ref = fix_f0(tw @ (seg * w), tw @ (seg * dw), f, fs, nfft, H)
            new = fix_f0(tw @ (seg * w0), tw @ (seg * dw0), f, fs, nfft, H)
            worst_f = max(worst_f, abs(new[0] - ref[0]) / abs(ref[0]))
            worst_s = max(worst_s, abs(new[1] - ref[1]) / abs(ref[1]))
    print(f"200 candidates x 7 frames: refined f0 differs by <= {worst_f:.1e} relative, score by <= {worst_s:.1e}")


if __name__ == "__main__":
    main()
