"""Shared parity checks; the same assertions run against the host emulation (CPU, -m "not gpu")
and against the CUDA library (-m gpu).  Tolerance: BASELINE.json's north_star -- 1e-6 relative
for f0 / spectrogram / aperiodicity, frame counts and time_axis bit-exact."""
import numpy as np

from refworld import rel_err

TOL = 1e-6


def to_np(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def make(world, a, dtype=np.float64):
    """host array -> the array type `world` works on"""
    a = np.ascontiguousarray(a, dtype=dtype)
    if world.xp == "torch":
        import torch
        return torch.from_numpy(a).to(f"cuda:{world.device}")
    return a


def wav_from_golden(golden):
    return golden["pcm"].astype(np.float64) / 32768.0, int(golden["fs"])


def assert_close(got, want, what, tol=TOL):
    r = rel_err(to_np(got), want)
    assert r.max() <= tol, f"{what}: max rel err {r.max():.3e}, {(r > tol).mean():.2%} of entries beyond {tol}"


def check_randn(world, golden):
    n = 1000008
    out = make(world, np.zeros(n, dtype=np.uint32), dtype=np.uint32)
    world.randn_stream(n, out)
    world.synchronize()
    vals = to_np(out).astype(np.float64) / 268435456.0 - 6.0
    assert np.array_equal(vals[:32], golden["randn_first32"])      # bit exact
    assert np.array_equal(vals[1000000:1000008], golden["randn_at_1e6"])


def check_golden_cheaptrick_d4c_stonemask(world, golden):
    x, fs = wav_from_golden(golden)
    xb = make(world, x[None, :])
    t = make(world, golden["time_axis"][None, :])
    f0 = make(world, golden["f0_stonemask"][None, :])
    fft = int(golden["fft_size"])
    opt = world.cheaptrick_option(fs)
    assert opt.fft_size == fft
    sp = world.cheaptrick(xb, fs, t, f0, opt)
    ap = world.d4c(xb, fs, t, f0, fft)
    sm = world.stonemask(xb, fs, t, make(world, golden["f0_dio"][None, :]))
    sm40 = world.stonemask(xb, fs, t, make(world, golden["f0_dio_floor40"][None, :]))
    world.synchronize()
    assert_close(sp[0], golden["sp"], "spectrogram (vaiueo2d, DIO path)")
    assert_close(ap[0], golden["ap"], "aperiodicity (vaiueo2d, DIO path)")
    assert_close(sm[0], golden["f0_stonemask"], "StoneMask f0")
    assert_close(sm40[0], golden["f0_stonemask_floor40"], "StoneMask f0 (floor 40)")
    # Harvest-path f0 through the same kernels (rows subsampled in the fixture)
    f0h = make(world, golden["f0_harvest"][None, :])
    sp_h = world.cheaptrick(xb, fs, t, f0h, opt)
    ap_h = world.d4c(xb, fs, t, f0h, fft)
    world.synchronize()
    assert_close(to_np(sp_h)[0][::4], golden["sp_harvest_rows"], "spectrogram (Harvest f0)")
    assert_close(to_np(ap_h)[0][::4], golden["ap_harvest_rows"], "aperiodicity (Harvest f0)")


def check_batch_vs_ref(world, ref, fs, n_samples, seeds, f0_method="dio", zero_tail=0, ragged=False,
                       stages=("f0", "sp", "ap")):
    """Synthetic ragged batch through the batched ABI vs the reference run utterance by utterance."""
    from synth import synth_batch
    x = synth_batch(seeds, fs, n_samples, device="cpu", zero_tail=zero_tail).numpy()
    n = len(seeds)
    lens = [n_samples - (37 * i * (fs // 100)) % (n_samples // 3) for i in range(n)] if ragged else [n_samples] * n
    xb = make(world, x)
    if f0_method == "dio":
        t, f0, fl = world.dio(xb, fs, x_lengths=lens)
        f0 = world.stonemask(xb, fs, t, f0, x_lengths=lens, f0_lengths=fl)
    elif f0_method == "harvest":
        t, f0, fl = world.harvest(xb, fs, x_lengths=lens)
    else:  # reference f0 handed in
        fl = [ref.frames(fs, l) for l in lens]
        tn = np.zeros((n, max(fl))); fn = np.zeros((n, max(fl)))
        for u in range(n):
            tr, fr = ref.dio(x[u, :lens[u]], fs)
            fn[u, :fl[u]] = ref.stonemask(x[u, :lens[u]], fs, tr, fr); tn[u, :fl[u]] = tr
        t, f0 = make(world, tn), make(world, fn)
    opt = world.cheaptrick_option(fs)
    sp = world.cheaptrick(xb, fs, t, f0, opt, x_lengths=lens, f0_lengths=fl) if "sp" in stages else None
    ap = world.d4c(xb, fs, t, f0, opt.fft_size, x_lengths=lens, f0_lengths=fl) if "ap" in stages else None
    world.synchronize()
    t, f0 = to_np(t), to_np(f0)
    flips = 0
    for u in range(n):
        xu = x[u, :lens[u]]
        if f0_method == "harvest":
            tr, fr = ref.harvest(xu, fs)
        else:
            tr, fr = ref.dio(xu, fs)
            fr = ref.stonemask(xu, fs, tr, fr)
        assert len(tr) == fl[u]
        assert np.array_equal(t[u, :fl[u]], tr), "time_axis must be bit exact"
        if "f0" in stages:
            flips += int(np.sum((f0[u, :fl[u]] == 0) != (fr == 0)))
            assert_close(f0[u, :fl[u]], fr, f"f0 utt {u}")
        # spectral stages are compared on the f0 the GPU path produced (== reference within 1e-6,
        # but window lengths derive from it, so hand the reference exactly the same values)
        fu = np.ascontiguousarray(f0[u, :fl[u]])
        if sp is not None:
            assert_close(to_np(sp)[u, :fl[u]], ref.cheaptrick(xu, fs, tr, fu, opt), f"spectrogram utt {u}")
        if ap is not None:
            assert_close(to_np(ap)[u, :fl[u]], ref.d4c(xu, fs, tr, fu, opt.fft_size), f"aperiodicity utt {u}")
    assert flips == 0


def check_golden_dio(world, golden):
    x, fs = wav_from_golden(golden)
    xb = make(world, x[None, :])
    t, f0, fl = world.dio(xb, fs)
    o = world.dio_option(); o.f0_floor = 40.0
    _, f0_40, _ = world.dio(xb, fs, o)
    world.synchronize()
    assert fl[0] == len(golden["time_axis"])
    assert np.array_equal(to_np(t)[0], golden["time_axis"])
    assert_close(f0[0], golden["f0_dio"], "DIO f0")
    assert_close(f0_40[0], golden["f0_dio_floor40"], "DIO f0 (floor 40)")


def check_golden_harvest(world, golden):
    x, fs = wav_from_golden(golden)
    xb = make(world, x[None, :])
    t, f0, fl = world.harvest(xb, fs)
    o = world.harvest_option(); o.f0_floor = 40.0
    _, f0_40, _ = world.harvest(xb, fs, o)
    world.synchronize()
    assert np.array_equal(to_np(t)[0], golden["time_axis"])
    assert_close(f0[0], golden["f0_harvest"], "Harvest f0")
    assert_close(f0_40[0], golden["f0_harvest_floor40"], "Harvest f0 (floor 40)")


def check_edge_cases(world, ref):
    """Silence, very short and ragged utterances, non-default frame periods."""
    from synth import synth_batch
    fs = 16000
    x = synth_batch([61, 62, 63], fs, 8000).numpy()
    x[1, :] = 0.0                      # digital silence: f0 = 0 everywhere, randn-only frames downstream
    lens = [8000, 8000, 900]           # third utterance is shorter than most analysis windows
    xb = make(world, x)
    for method in ("dio", "harvest"):
        for fp in (5.0, 10.0, 2.5):
            if method == "dio":
                o = world.dio_option(); o.frame_period = fp
                ro = ref.dio_option(); ro.frame_period = fp
                t, f0, fl = world.dio(xb, fs, o, x_lengths=lens)
            else:
                o = world.harvest_option(); o.frame_period = fp
                ro = ref.harvest_option(); ro.frame_period = fp
                t, f0, fl = world.harvest(xb, fs, o, x_lengths=lens)
            world.synchronize()
            for u in range(3):
                xu = x[u, :lens[u]]
                tr, fr = (ref.dio(xu, fs, ro) if method == "dio" else ref.harvest(xu, fs, ro))
                assert fl[u] == len(tr)
                assert np.array_equal(to_np(t)[u, :fl[u]], tr)
                assert_close(to_np(f0)[u, :fl[u]], fr, f"{method} fp={fp} utt {u}")
    # spectral stages on the silent and the short utterance (reference f0)
    fl = [ref.frames(fs, l) for l in lens]
    tn = np.zeros((3, max(fl))); fn = np.zeros((3, max(fl)))
    for u in range(3):
        tr, fr = ref.dio(x[u, :lens[u]], fs)
        tn[u, :fl[u]] = tr; fn[u, :fl[u]] = ref.stonemask(x[u, :lens[u]], fs, tr, fr)
    t, f0 = make(world, tn), make(world, fn)
    opt = world.cheaptrick_option(fs)
    sp = world.cheaptrick(xb, fs, t, f0, opt, x_lengths=lens, f0_lengths=fl)
    ap = world.d4c(xb, fs, t, f0, opt.fft_size, x_lengths=lens, f0_lengths=fl)
    world.synchronize()
    for u in range(3):
        xu = x[u, :lens[u]]
        assert_close(to_np(sp)[u, :fl[u]], ref.cheaptrick(xu, fs, tn[u, :fl[u]], fn[u, :fl[u]], opt), f"sp edge utt {u}")
        assert_close(to_np(ap)[u, :fl[u]], ref.d4c(xu, fs, tn[u, :fl[u]], fn[u, :fl[u]], opt.fft_size), f"ap edge utt {u}")


def check_synthesis(world, ref, golden):
    """batched Synthesis() vs the reference (relative to the waveform peak)."""
    from synth import synth_batch
    x, fs = wav_from_golden(golden)
    fft = int(golden["fft_size"])
    f0 = golden["f0_stonemask"]; sp = golden["sp"]; ap = golden["ap"]
    yr = ref.synthesis(f0, sp, ap, fft, 5.0, fs, len(x))
    y = world.synthesis(make(world, f0[None]), make(world, sp[None]), make(world, ap[None]), fft, 5.0, fs, len(x))
    world.synchronize()
    assert np.abs(to_np(y)[0] - yr).max() <= 1e-9 * np.abs(yr).max()
    # ragged synthetic batch, parameters from the reference analysis
    fs2, n = 16000, 12000
    xs = synth_batch([81, 82], fs2, n).numpy()
    lens = [12000, 9000]
    L = [ref.frames(fs2, l) for l in lens]
    opt = ref.cheaptrick_option(fs2)
    bins = opt.fft_size // 2 + 1
    F = np.zeros((2, max(L))); S = np.ones((2, max(L), bins)); A = np.ones((2, max(L), bins))
    refs = []
    for u in range(2):
        xu = xs[u, :lens[u]]
        t, f = ref.harvest(xu, fs2)
        s_ = ref.cheaptrick(xu, fs2, t, f, opt); a_ = ref.d4c(xu, fs2, t, f, opt.fft_size)
        F[u, :L[u]] = f; S[u, :L[u]] = s_; A[u, :L[u]] = a_
        refs.append(ref.synthesis(f, s_, a_, opt.fft_size, 5.0, fs2, lens[u]))
    y = world.synthesis(make(world, F), make(world, S), make(world, A), opt.fft_size, 5.0, fs2, n, f0_lengths=L,
                        y_lengths=lens)
    world.synchronize()
    for u in range(2):
        assert np.abs(to_np(y)[u, :lens[u]] - refs[u]).max() <= 1e-9 * np.abs(refs[u]).max(), f"synthesis utt {u}"


def check_fft_known_answers(world):
    """The shared-memory FFT against numpy's (conventions of SURVEY.md App. A0), every size it serves."""
    rng = np.random.RandomState(7)
    for lg in range(2, 14):
        n = 1 << lg
        x = rng.standard_normal(n)
        out = make(world, np.zeros(n + 2))
        world.rfft_test(make(world, x), out)
        world.synchronize()
        got = to_np(out).reshape(-1, 2)
        want = np.fft.rfft(x)
        err = np.abs(got[:, 0] + 1j * got[:, 1] - want).max() / np.abs(want).max()
        assert err < 1e-13, f"rfft n={n}: {err:.2e}"
