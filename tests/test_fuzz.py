"""Short randomised parity sweeps under pytest (the long campaigns are run by hand: tests/fuzz/README.md).
Random rates 8-48 kHz, lengths, options of both F0 estimators, signal kinds speech / noise / tone / impulses /
clipped / DC / embedded digital silence; every case is compared with the compiled reference (1e-6, no V/UV flip,
frame counts and time axis bit exact), spectral stages on the reference's f0 included.  The CPU test runs the kernel
sources as host emulation; the -m gpu test runs the CUDA library through the C ABI on cuda:0."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz"))


def _run(world, ref, n_cases, seed):
    import fuzz_emu_parity as fz
    lines = []
    bad = fz.run(world, ref, n_cases, seed, out=lines.append)
    assert bad == 0, "\n".join(l for l in lines if "MISMATCH" in l or "ERROR" in l)


def test_fuzz_f0_and_spectral_emu(emu, ref):
    _run(emu, ref, 16, 1)


@pytest.mark.gpu
def test_gpu_fuzz_f0_and_spectral(gpu_world, ref):
    _run(gpu_world, ref, 48, 1)
    _run(gpu_world, ref, 24, 2)
