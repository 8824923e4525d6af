"""Randomised RAGGED BATCHES through the F0 estimators and the one-call host pipeline (kernel sources on the host
vs the compiled reference run utterance by utterance): batch composition must not matter.  CPU only.
Usage: python tests/fuzz/fuzz_emu_batches.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refworld import RefWorld, rel_err  # noqa: E402
from world_b200.api import World, WorldError, F0_HARVEST, F0_DIO_STONEMASK  # noqa: E402
from synth import synth_batch  # noqa: E402

TOL = 1e-6


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ref = RefWorld()
    emu = World(lib_path=os.environ.get("WB_EMU_LIB", os.path.join(ROOT, "tests", "emu", "libworld_b200_emu.so")), array_module="numpy")
    bad = 0
    for case in range(n_cases):
        fs = int(rng.choice([8000, 16000, 22050, 48000]))
        nu = int(rng.integers(2, 5))
        n = int(rng.uniform(0.2, 0.6) * fs)
        x = synth_batch([int(s) for s in rng.integers(1, 1 << 30, size=nu)], fs, n).numpy()
        lens = [n] + [int(rng.integers(max(2, n // 20), n + 1)) for _ in range(nu - 1)]
        if case % 3 == 0:
            x[1, : lens[1] // 2] = 0.0
        method = F0_HARVEST if case % 2 else F0_DIO_STONEMASK
        os.environ["WB_HOST_SUB"], os.environ["WB_HOST_CHUNK"] = str(int(rng.integers(1, 3))), str(int(rng.integers(1, 4)))
        t0 = time.time()
        try:
            opt = emu.analysis_option(fs, method)
            ta, fa, spa, apa, fl = emu.analyze_host(x, fs, opt, x_lengths=lens)
            worst = 0.0
            for u in range(nu):
                xu = np.ascontiguousarray(x[u, :lens[u]])
                if method == F0_HARVEST:
                    tr, fr = ref.harvest(xu, fs)
                else:
                    tr, fr = ref.dio(xu, fs)
                    fr = ref.stonemask(xu, fs, tr, fr)
                assert fl[u] == len(tr) and np.array_equal(ta[u, :fl[u]], tr), "time axis"
                got = fa[u, :fl[u]]
                assert not ((got > 0) != (fr > 0)).any(), f"V/UV flips in utterance {u}"
                worst = max(worst, rel_err(got, fr).max())
                assert not fa[u, fl[u]:].any() and not spa[u, fl[u]:].any(), "padding not zero"
                if fs >= 16000:
                    fu = np.ascontiguousarray(got)
                    worst = max(worst, rel_err(spa[u, :fl[u]], ref.cheaptrick(xu, fs, tr, fu)).max(),
                                rel_err(apa[u, :fl[u]], ref.d4c(xu, fs, tr, fu, opt.cheaptrick.fft_size)).max())
            assert worst <= TOL, f"max rel err {worst:.1e}"
            status = f"ok ({worst:.1e})"
        except (AssertionError, WorldError) as e:
            status = f"FAIL {e}"
            bad += 1
        print(f"case {case:3d} fs {fs:5d} utts {nu} lens {lens} method {'harvest' if method == F0_HARVEST else 'dio'}  {status}  ({time.time() - t0:.1f}s)", flush=True)
    print(f"{n_cases - bad}/{n_cases} ragged batches agree within {TOL}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
