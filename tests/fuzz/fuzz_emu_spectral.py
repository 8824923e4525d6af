"""Randomised parity sweep of StoneMask / CheapTrick / D4C (kernel sources on the host vs the compiled
reference) with adversarial f0 contours: values below the floors, near fs/2, jumps, zeros, time axes with
non-default frame periods, ragged batches.  CPU only.  Usage: python tests/fuzz/fuzz_emu_spectral.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refworld import RefWorld, rel_err  # noqa: E402
from world_b200.api import World, WorldError  # noqa: E402
from synth import synth_batch  # noqa: E402

TOL = 1e-6


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ref = RefWorld()
    emu = World(lib_path=os.environ.get("WB_EMU_LIB", os.path.join(ROOT, "tests", "emu", "libworld_b200_emu.so")), array_module="numpy")
    bad = 0
    for case in range(n_cases):
        fs = int(rng.choice([16000, 22050, 32000, 44100, 48000]))
        nu = int(rng.integers(1, 4))
        n = int(rng.uniform(0.1, 0.5) * fs)
        fp = float(rng.choice([1.0, 2.5, 5.0, 5.0, 10.0]))
        x = synth_batch([int(s) for s in rng.integers(1, 1 << 30, size=nu)], fs, n).numpy()
        if case % 5 == 0:
            x[0] = rng.normal(size=n) * 0.05
        if case % 7 == 0:
            x[-1, n // 3:] = 0.0
        lens = [n - int(rng.integers(0, n // 2)) if u else n for u in range(nu)]
        fl = [ref.frames(fs, l, fp) for l in lens]
        L = max(fl)
        t = np.zeros((nu, L)); f0 = np.zeros((nu, L))
        co = emu.cheaptrick_option(fs); rco = ref.cheaptrick_option(fs)
        if case % 4 == 1:
            for q in (co, rco):
                q.f0_floor = 40.0
                q.fft_size = ref.lib.GetFFTSizeForCheapTrick(fs, __import__("ctypes").byref(q))
        if case % 6 == 2:
            co.q1 = rco.q1 = -0.2
        floor_ct = 3.0 * fs / (co.fft_size - 3.0)
        for u in range(nu):
            t[u, :fl[u]] = np.arange(fl[u]) * fp / 1000.0
            kind = rng.choice(["smooth", "random", "edges"])
            if kind == "smooth":
                f = rng.uniform(60, 500) * (1 + 0.3 * np.sin(np.arange(fl[u]) * rng.uniform(0.01, 0.2)))
            elif kind == "random":
                f = rng.uniform(floor_ct * 1.01, min(1500.0, fs / 4.0), size=fl[u])
            else:
                f = rng.choice([0.0, floor_ct * 1.0001, floor_ct * 0.5, 47.0, 46.9, 71.0, 800.0, fs / 6.0], size=fl[u])
            f[rng.uniform(size=fl[u]) < 0.2] = 0.0
            f0[u, :fl[u]] = f
        do = emu.d4c_option(); rdo = ref.d4c_option()
        if case % 3 == 2:
            do.threshold = rdo.threshold = float(rng.choice([0.0, 0.5, 0.95]))
        t0 = time.time()
        msg = ""
        try:
            sm = emu.stonemask(x, fs, t, f0, x_lengths=lens, f0_lengths=fl)
            sp = emu.cheaptrick(x, fs, t, f0, co, x_lengths=lens, f0_lengths=fl)
            ap = emu.d4c(x, fs, t, f0, co.fft_size, do, x_lengths=lens, f0_lengths=fl)
            emu.synchronize()
            worst = [0.0, 0.0, 0.0]
            for u in range(nu):
                xu = np.ascontiguousarray(x[u, :lens[u]]); tu = np.ascontiguousarray(t[u, :fl[u]]); fu = np.ascontiguousarray(f0[u, :fl[u]])
                worst[0] = max(worst[0], rel_err(sm[u, :fl[u]], ref.stonemask(xu, fs, tu, fu)).max())
                worst[1] = max(worst[1], rel_err(sp[u, :fl[u]], ref.cheaptrick(xu, fs, tu, fu, rco)).max())
                worst[2] = max(worst[2], rel_err(ap[u, :fl[u]], ref.d4c(xu, fs, tu, fu, rco.fft_size, rdo)).max())
                assert not sp[u, fl[u]:].any() and not ap[u, fl[u]:].any(), "padded frames written"
            msg = f"stonemask {worst[0]:.1e} sp {worst[1]:.1e} ap {worst[2]:.1e}"
            assert max(worst) <= TOL, "mismatch"
            status = "ok"
        except AssertionError as e:
            status = f"MISMATCH {e}"
            bad += 1
        except WorldError as e:
            status = f"ERROR {e}"
            bad += 1
        print(f"case {case:3d} fs {fs:5d} utts {nu} n {n:6d} fp {fp:4.1f} fft {co.fft_size:5d}  {msg:52s} {status}  ({time.time() - t0:.1f}s)", flush=True)
    print(f"{n_cases - bad}/{n_cases} cases agree within {TOL}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
