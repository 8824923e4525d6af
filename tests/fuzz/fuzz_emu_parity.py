"""Randomised parity sweep against the compiled reference: the kernel sources as host emulation (default, CPU only)
or the CUDA library on cuda:0 (WB_FUZZ_GPU=1; tests/test_fuzz.py runs a short sweep of each under pytest).
Usage: python tests/fuzz/fuzz_emu_parity.py [n_cases] [seed]   -- prints one line per case, exits non-zero on a mismatch."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refworld import RefWorld, rel_err  # noqa: E402
from world_b200.api import World  # noqa: E402
from synth import synth_batch  # noqa: E402
from test_parity_common import make, to_np  # noqa: E402

TOL = 1e-6


def make_signal(rng, fs, n):
    kind = rng.choice(["speech", "speech", "speech", "noise", "tone", "silence_mix", "dc", "clipped", "impulses"])
    x = synth_batch([int(rng.integers(1, 1 << 30))], fs, n).numpy()[0]
    t = np.arange(n) / fs
    if kind == "noise":
        x = rng.normal(size=n) * 0.1
    elif kind == "tone":
        x = 0.3 * np.sin(2 * np.pi * rng.uniform(60, 600) * t) + 1e-4 * rng.normal(size=n)
    elif kind == "silence_mix":
        a, b = sorted(rng.integers(0, n, size=2))
        x[a:b] = 0.0
    elif kind == "dc":
        x = x + rng.uniform(-0.3, 0.3)
    elif kind == "clipped":
        x = np.clip(x * 8, -1, 1)
    elif kind == "impulses":
        x = np.zeros(n)
        x[::max(1, int(fs / rng.uniform(80, 300)))] = 0.5
        x += 1e-5 * rng.normal(size=n)
    return kind, np.ascontiguousarray(x)


def open_world():
    if os.environ.get("WB_FUZZ_GPU"):
        import torch
        torch.cuda.set_device(0)
        return World(device=0)
    return World(lib_path=os.environ.get("WB_EMU_LIB", os.path.join(ROOT, "tests", "emu", "libworld_b200_emu.so")), array_module="numpy")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    bad = run(open_world(), RefWorld(), n_cases, seed)
    sys.exit(1 if bad else 0)


def run(emu, ref, n_cases, seed, out=print):
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        fs = int(rng.choice([8000, 11025, 16000, 22050, 32000, 44100, 48000]))
        n = int(rng.uniform(0.15, 0.9) * fs)
        kind, x = make_signal(rng, fs, n)
        fp = float(rng.choice([1.0, 2.5, 5.0, 5.0, 10.0, 3.7]))
        method = rng.choice(["dio", "harvest"])
        t0 = time.time()
        msg = []
        try:
            if method == "dio":
                o = emu.dio_option(); ro = ref.dio_option()
                for q in (o, ro):
                    q.frame_period = fp
                    q.speed = int(1 + case % 12)
                    q.f0_floor = [71.0, 40.0, 90.0][case % 3]
                    q.f0_ceil = [800.0, 400.0, 1000.0][(case // 3) % 3]
                    q.channels_in_octave = [2.0, 4.0, 1.0][(case // 2) % 3]
                    q.allowed_range = [0.1, 0.2][case % 2]
                xd = make(emu, x[None])
                t, f0, fl = emu.dio(xd, fs, o)
                emu.synchronize()
                tr, fr = ref.dio(x, fs, ro)
                # same input to both (the reference's DIO f0): StoneMask rounds f0 * fft / fs * k to a bin, so
                # inputs that differ in the last digits can legitimately land on different bins
                f0s = to_np(emu.stonemask(xd, fs, make(emu, tr[None]), make(emu, np.ascontiguousarray(fr)[None]))); emu.synchronize()
                frs = ref.stonemask(x, fs, tr, fr)
                e_sm = rel_err(f0s[0], frs).max()
                msg.append(f"stonemask {e_sm:.1e}")
                assert e_sm <= TOL
            else:
                o = emu.harvest_option(); ro = ref.harvest_option()
                for q in (o, ro):
                    q.frame_period = fp
                    q.f0_floor = [71.0, 40.0, 100.0][case % 3]
                    q.f0_ceil = [800.0, 500.0, 1100.0][(case // 3) % 3]
                xd = make(emu, x[None])
                t, f0, fl = emu.harvest(xd, fs, o)
                emu.synchronize()
                tr, fr = ref.harvest(x, fs, ro)
            t, f0 = to_np(t), to_np(f0)
            assert fl[0] == len(tr) and np.array_equal(t[0], tr), "time axis"
            flips = int(((f0[0] > 0) != (fr > 0)).sum())
            e_f0 = rel_err(f0[0], fr).max() if flips == 0 else float("inf")
            msg.append(f"f0 {e_f0:.1e} voiced {int((fr > 0).sum())}/{len(fr)}")
            assert flips == 0 and e_f0 <= TOL, f"f0 mismatch, {flips} V/UV flips"
            # spectral stages on the reference's f0, non-default CheapTrick / D4C options now and then
            if fs >= 16000:
                co = emu.cheaptrick_option(fs); rco = ref.cheaptrick_option(fs)
                if case % 4 == 1:
                    for q in (co, rco):
                        q.q1 = -0.09
                        q.fft_size = q.fft_size * 2
                do = emu.d4c_option(); rdo = ref.d4c_option()
                if case % 5 == 2:
                    do.threshold = rdo.threshold = 0.5
                frc = np.ascontiguousarray(fr)
                sp = emu.cheaptrick(xd, fs, make(emu, tr[None]), make(emu, frc[None]), co)
                ap = emu.d4c(xd, fs, make(emu, tr[None]), make(emu, frc[None]), co.fft_size, do)
                emu.synchronize()
                sp, ap = to_np(sp), to_np(ap)
                spr, apr = ref.cheaptrick(x, fs, tr, frc, rco), ref.d4c(x, fs, tr, frc, rco.fft_size, rdo)
                e_sp = rel_err(sp[0], spr).max()
                e_ap = rel_err(ap[0], apr).max()
                msg.append(f"sp {e_sp:.1e} ap {e_ap:.1e}")
                assert e_sp <= TOL and e_ap <= TOL
                # the same frames through the fused kernels: coded rows against the reference's two-call chain.
                # Coded values live in the log domain -- a relative error e of a bin is an ABSOLUTE error of
                # 20 / ln(10) * e dB in a coded aperiodicity and of at most sqrt(fft_size / 2) * e in a cepstral
                # coefficient (orthonormal DCT of the log envelope) -- so the 1e-6 relative bound of the rows becomes
                # these absolute bounds (near-periodic tones reach ap errors of 5e-7, i.e. 4e-6 dB around -2 dB).
                dims = [24, 60, 1, co.fft_size // 4 + 1][case % 4]
                csp = emu.cheaptrick_coded(xd, fs, make(emu, tr[None]), make(emu, frc[None]), dims, co)
                cap = emu.d4c_coded(xd, fs, make(emu, tr[None]), make(emu, frc[None]), co.fft_size, do)
                emu.synchronize()
                want = ref.code_spectral_envelope(spr, fs, rco.fft_size, dims)
                e_csp = np.abs(to_np(csp)[0] - want).max()
                want = ref.code_aperiodicity(apr, fs, rco.fft_size)
                e_cap = np.abs(to_np(cap)[0] - want).max()
                msg.append(f"coded abs {e_csp:.1e} {e_cap:.1e}")
                assert e_csp <= np.sqrt(rco.fft_size / 2.0) * TOL and e_cap <= 20.0 / np.log(10.0) * TOL
            status = "ok"
        except AssertionError as e:
            status = f"MISMATCH {e}"
            bad += 1
        except Exception as e:   # library errors (EDOMAIN etc.) are reported, not hidden
            status = f"ERROR {type(e).__name__}: {e}"
            bad += 1
        out(f"case {case:3d} fs {fs:5d} n {n:6d} {kind:11s} {method:7s} fp {fp:4.1f}  {'  '.join(msg):84s} {status}  ({time.time() - t0:.1f}s)")
    out(f"{n_cases - bad}/{n_cases} cases agree with the reference within {TOL}; {bad} failures")
    return bad


if __name__ == "__main__":
    main()
