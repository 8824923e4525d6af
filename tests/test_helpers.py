"""The public MATLAB-style helpers of world/matlabfunctions.h (host only) against the reference's own: bit-identical
on random inputs, including the corner cases of histc (edges below the first node, on nodes, beyond the last)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from world_b200 import api
    return api.load_library()


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def bind(L):
    P = C.c_void_p
    L.fftshift.argtypes = [P, C.c_int, P]
    L.histc.argtypes = [P, C.c_int, P, C.c_int, P]
    L.interp1.argtypes = [P, P, C.c_int, P, C.c_int, P]
    L.decimate.argtypes = [P, C.c_int, C.c_int, P]
    L.matlab_round.argtypes = [C.c_double]; L.matlab_round.restype = C.c_int
    L.diff.argtypes = [P, C.c_int, P]
    L.interp1Q.argtypes = [C.c_double, C.c_double, P, C.c_int, P, C.c_int, P]
    L.randn.argtypes = [P]; L.randn.restype = C.c_double
    L.randn_reseed.argtypes = [P]
    L.matlab_std.argtypes = [P, C.c_int]; L.matlab_std.restype = C.c_double
    for f in (L.fftshift, L.histc, L.interp1, L.decimate, L.diff, L.interp1Q, L.randn_reseed):
        f.restype = None
    return L


def test_helpers_are_bit_identical_to_the_reference(lib, ref):
    A, B = bind(lib), bind(ref.lib)
    rng = np.random.default_rng(7)
    for trial in range(200):
        nx = int(rng.integers(2, 40))
        x = np.sort(rng.normal(size=nx)).copy()
        if trial % 5 == 0:
            x = np.arange(nx, dtype=np.float64)                      # edges can fall exactly on nodes
        y = rng.normal(size=nx)
        ne = int(rng.integers(1, 60))
        lo, hi = (x[0] - 1, x[-1] + 1) if trial % 3 else (x[0], x[-1])
        xi = np.sort(rng.uniform(lo, hi, size=ne)).copy()
        if trial % 5 == 0:
            xi = np.sort(np.round(xi * 2) / 2).copy()
        ia = np.zeros(ne, dtype=np.int32); ib = np.zeros(ne, dtype=np.int32)
        A.histc(ptr(x), nx, ptr(xi), ne, ptr(ia)); B.histc(ptr(x), nx, ptr(xi), ne, ptr(ib))
        assert np.array_equal(ia, ib), (trial, x, xi, ia, ib)
        ya = np.zeros(ne); yb = np.zeros(ne)
        A.interp1(ptr(x), ptr(y), nx, ptr(xi), ne, ptr(ya)); B.interp1(ptr(x), ptr(y), nx, ptr(xi), ne, ptr(yb))
        assert np.array_equal(ya, yb)
        # interp1Q inside the grid
        x0, dx = float(rng.normal()), float(rng.uniform(0.1, 2.0))
        q = np.sort(rng.uniform(x0, x0 + dx * (nx - 1), size=ne)).copy()
        A.interp1Q(x0, dx, ptr(y), nx, ptr(q), ne, ptr(ya)); B.interp1Q(x0, dx, ptr(y), nx, ptr(q), ne, ptr(yb))
        assert np.array_equal(ya, yb)
        da = np.zeros(nx); db = np.zeros(nx)
        A.diff(ptr(y), nx, ptr(da)); B.diff(ptr(y), nx, ptr(db))
        assert np.array_equal(da, db)
        assert A.matlab_std(ptr(y), nx) == B.matlab_std(ptr(y), nx)
        v = float(rng.normal() * 100)
        assert A.matlab_round(v) == B.matlab_round(v) and A.matlab_round(0.5) == 1 and A.matlab_round(-0.5) == -1
        n2 = 2 * int(rng.integers(1, 30))
        z = rng.normal(size=n2); za = np.zeros(n2); zb = np.zeros(n2)
        A.fftshift(ptr(z), n2, ptr(za)); B.fftshift(ptr(z), n2, ptr(zb))
        assert np.array_equal(za, zb)
    for r in range(1, 14):                                            # 1 and 13: the all-zero default branch
        n = int(rng.integers(40, 3000))
        x = rng.normal(size=n)
        ya = np.full(n + 16, np.nan); yb = np.full(n + 16, np.nan)
        A.decimate(ptr(x), n, r, ptr(ya)); B.decimate(ptr(x), n, r, ptr(yb))
        assert np.array_equal(np.isnan(ya), np.isnan(yb)) and np.array_equal(ya[~np.isnan(ya)], yb[~np.isnan(yb)]), r
    sa = (C.c_uint32 * 4)(); sb = (C.c_uint32 * 4)()
    A.randn_reseed(sa); B.randn_reseed(sb)
    assert [A.randn(sa) for _ in range(1000)] == [B.randn(sb) for _ in range(1000)] and list(sa) == list(sb)
