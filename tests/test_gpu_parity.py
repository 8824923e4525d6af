"""Parity tests proper: the CUDA library (through the C ABI) against the golden vectors and the
compiled reference, on a real B200.  Tolerance 1e-6 relative (north_star); integers bit exact."""
import numpy as np
import pytest

import test_parity_common as pc

pytestmark = pytest.mark.gpu


def test_gpu_randn_stream(gpu_world, golden):
    pc.check_randn(gpu_world, golden)


def test_gpu_golden_cheaptrick_d4c_stonemask(gpu_world, golden):
    pc.check_golden_cheaptrick_d4c_stonemask(gpu_world, golden)


@pytest.mark.parametrize("fs,n,seeds", [(16000, 16000, [1, 2, 3, 4, 5]), (48000, 24000, [6, 7]), (22050, 11025, [8])])
def test_gpu_spectral_stages_on_reference_f0(gpu_world, ref, fs, n, seeds):
    pc.check_batch_vs_ref(gpu_world, ref, fs, n, seeds, f0_method="ref", ragged=len(seeds) > 1, stages=("sp", "ap"))


def test_gpu_zero_tail(gpu_world, ref):
    pc.check_batch_vs_ref(gpu_world, ref, 16000, 16000, [9], f0_method="ref", zero_tail=6000, stages=("sp", "ap"))


def test_gpu_golden_dio(gpu_world, golden):
    pc.check_golden_dio(gpu_world, golden)


@pytest.mark.parametrize("fs,n,seeds", [(16000, 48000, [11, 12, 13, 14]), (48000, 48000, [15, 16]), (22050, 22050, [17])])
def test_gpu_dio_path_end_to_end(gpu_world, ref, fs, n, seeds):
    pc.check_batch_vs_ref(gpu_world, ref, fs, n, seeds, f0_method="dio", ragged=len(seeds) > 1)


def test_gpu_dio_decimated(gpu_world, ref):
    import torch
    from synth import synth_batch
    x = synth_batch([21], 44100, 44100)
    o = gpu_world.dio_option(); o.speed = 11
    ro = ref.dio_option(); ro.speed = 11
    t, f0, fl = gpu_world.dio(x.cuda(), 44100, o)
    gpu_world.synchronize()
    tr, fr = ref.dio(x[0].numpy(), 44100, ro)
    assert np.array_equal(t[0].cpu().numpy(), tr)
    pc.assert_close(f0[0], fr, "DIO speed=11")


def test_gpu_golden_harvest(gpu_world, golden):
    pc.check_golden_harvest(gpu_world, golden)


@pytest.mark.parametrize("fs,n,seeds", [(16000, 48000, [31, 32, 33, 34]), (48000, 48000, [35, 36]), (22050, 22050, [37])])
def test_gpu_harvest_path_end_to_end(gpu_world, ref, fs, n, seeds):
    pc.check_batch_vs_ref(gpu_world, ref, fs, n, seeds, f0_method="harvest", ragged=len(seeds) > 1)
