"""Randomised parity sweep of batched Synthesis (kernel sources on the host vs the compiled reference): random
rates, frame periods, output lengths, f0 contours (incl. unvoiced stretches, values below the synthesis floor),
envelopes from the reference's own analysis.  CPU only.  Usage: python tests/fuzz/fuzz_emu_synthesis.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refworld import RefWorld  # noqa: E402
from world_b200.api import World, WorldError  # noqa: E402
from synth import synth_batch  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ref = RefWorld()
    emu = World(lib_path=os.environ.get("WB_EMU_LIB", os.path.join(ROOT, "tests", "emu", "libworld_b200_emu.so")), array_module="numpy")
    bad = 0
    for case in range(n_cases):
        fs = int(rng.choice([16000, 22050, 44100, 48000]))
        n = int(rng.uniform(0.15, 0.5) * fs)
        fp = float(rng.choice([2.5, 5.0, 5.0, 10.0]))
        x = synth_batch([int(rng.integers(1, 1 << 30))], fs, n).numpy()[0]
        o = ref.dio_option(); o.frame_period = fp
        t, f0 = ref.dio(x, fs, o)
        f0 = ref.stonemask(x, fs, t, f0)
        co = ref.cheaptrick_option(fs)
        sp = ref.cheaptrick(x, fs, t, f0, co)
        ap = ref.d4c(x, fs, t, f0, co.fft_size)
        f0s = f0.copy()
        if case % 3 == 1:
            f0s = f0s * rng.uniform(0.5, 2.0)                       # pitch shift
        if case % 4 == 2:
            f0s[rng.uniform(size=len(f0s)) < 0.3] = 0.0             # holes
        if case % 5 == 3:
            f0s[::7] = fs / co.fft_size * 0.5                       # below the synthesis floor
        y_len = int(n * rng.uniform(0.6, 1.3))
        t0 = time.time()
        try:
            y = emu.synthesis(f0s[None], sp[None], ap[None], co.fft_size, fp, fs, y_len)
            emu.synchronize()
            yr = ref.synthesis(f0s, sp, ap, co.fft_size, fp, fs, y_len)
            err = np.abs(y[0] - yr).max() / max(np.abs(yr).max(), 1e-300)
            status = "ok" if err <= 1e-9 else "MISMATCH"
            bad += status != "ok"
            msg = f"max err / peak {err:.1e}  peak {np.abs(yr).max():.3f}"
        except WorldError as e:
            status, msg = f"ERROR {e}", ""
            bad += 1
        print(f"case {case:3d} fs {fs:5d} n {n:6d} y {y_len:6d} fp {fp:4.1f} fft {co.fft_size:5d}  {msg:40s} {status}  ({time.time() - t0:.1f}s)", flush=True)
    print(f"{n_cases - bad}/{n_cases} cases agree within 1e-9 of the waveform peak")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
