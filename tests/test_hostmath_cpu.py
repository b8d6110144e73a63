"""CPU: the FP64 recurrences of sfa_numpy_b200/csrc/sfa_math.cuh (the SAME source the CUDA kernels
compile), built for the host by tests/hostmath, against the golden vectors of the reference and
oracle sweeps at the GPU tolerance (1e-10 relative + 1e-15 * column max)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import assert_close
from oracle import sfa_oracle as O

TOL = dict(rtol=1e-10, floor=1e-15)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_hostmath())
    vp, ci, cl, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double
    lib.hm_sph_radial.argtypes = [ci, cl, vp, cd, cd, ci, ci, vp]
    lib.hm_circ_radial.argtypes = [ci, cl, vp, cd, cd, ci, ci, vp]
    lib.hm_regularize.argtypes = [cl, vp, cd, ci, vp, vp]
    lib.hm_sht.argtypes = [ci, cl, vp, vp, ci, vp, vp]
    lib.hm_regularize_inverse.argtypes = [cl, vp, cd, ci, vp, vp]
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def radial(hm, fam, N, k, r, rs, setup, scale):
    k = np.ascontiguousarray(np.atleast_1d(k), dtype=float)
    out = np.empty((len(k), N + 1 if fam == 0 else 2 * N + 1), complex)
    (hm.hm_sph_radial if fam == 0 else hm.hm_circ_radial)(N, len(k), P(k), r, rs, setup, scale, P(out))
    return out


def sht(hm, N, azi, colat, w=None):
    azi = np.ascontiguousarray(np.atleast_1d(azi), dtype=float)
    colat = np.ascontiguousarray(np.atleast_1d(colat), dtype=float)
    Y = np.empty((len(azi), (N + 1) ** 2), complex)
    wp = P(np.ascontiguousarray(w, dtype=float)) if w is not None else None
    hm.hm_sht(N, len(azi), P(azi), P(colat), int(len(colat) == 1 and len(azi) > 1), wp, P(Y))
    return Y


def test_sh_recurrence(hm, golden):
    g = golden["angular"]
    assert_close(sht(hm, 4, g["ea4_azi"], g["ea4_colat"], g["ea4_w"]), g["ea4_Y"], **TOL)
    assert_close(sht(hm, 8, g["r8_azi"], g["r8_colat"]), g["r8_Y"], **TOL)
    assert_close(sht(hm, 8, g["eq8_azi"], np.pi / 2), g["eq8_Y"], **TOL)
    assert_close(sht(hm, 32, g["r32_azi"], g["r32_colat"]), g["r32_Y"], **TOL)
    az, co, w = O.fibonacci_grid(1500)
    for N in (8, 16, 32):
        assert_close(sht(hm, N, az, co, w), O.sht_matrix(N, az, co, w), **TOL, name="fib%d" % N)


@pytest.mark.parametrize("si,setup", list(enumerate(O.SETUPS)))
def test_radial_recurrences(hm, golden, si, setup):
    g = golden["radial"]
    assert_close(radial(hm, 0, 8, g["kr8"], 1.0, 0, si, 0), g["w8_" + setup], **TOL)
    assert_close(radial(hm, 0, 32, g["kr32"], 1.0, 0, si, 0), g["w32_" + setup], **TOL)
    assert_close(radial(hm, 1, 16, g["kr8"], 1.0, 0, si, 0), g["cw16_" + setup], **TOL)
    assert_close(radial(hm, 1, 32, g["kr32"], 1.0, 0, si, 0), g["cw32_" + setup], **TOL)
    k = g["k40"]
    assert_close(radial(hm, 0, 4, k, 0.042, 0, si, 1), g["pw4nr_" + setup], **TOL)
    assert_close(radial(hm, 0, 4, k, 0.042, 1.5, si, 2), O.spherical_ps(4, k, 0.042, 1.5, setup, replace_zeros=False), **TOL)
    assert_close(radial(hm, 1, 16, k, 0.1, 0, si, 1), O.circular_pw(16, k, 0.1, setup, replace_zeros=False), **TOL)
    assert_close(radial(hm, 1, 6, k[1:], 0.1, 2.0, si, 2), O.circular_ls(6, k[1:], 0.1, 2.0, setup, replace_zeros=False), **TOL)
    for N, kr in ((16, np.geomspace(0.01, 50, 512)), (32, np.geomspace(1e-6, 60, 300)), (50, np.linspace(0, 20, 100))):
        with np.errstate(all="ignore"):
            assert_close(radial(hm, 0, N, kr, 1.0, 0, si, 0), O.weights_2d(N, kr, setup), **TOL, name="sph%d" % N)
            assert_close(radial(hm, 1, N, kr, 1.0, 0, si, 0), O.circ_radial_weights_2d(N, kr, setup), **TOL, name="circ%d" % N)


@pytest.mark.parametrize("mi,method", list(enumerate(O.METHODS)))
def test_regularize(hm, golden, mi, method):
    g = golden["radial"]
    for key, a0 in (("reg_dn", 100.0), ("reg_big", 10.0)):
        if key == "reg_big" and method == "none":
            continue
        dn = np.ascontiguousarray(g[key]).reshape(-1)
        d, h = np.empty_like(dn), np.empty(dn.shape)
        hm.hm_regularize(dn.size, P(dn), a0, mi, P(d), P(h))
        pre = "reg" if key == "reg_dn" else "regbig"
        assert_close(d.reshape(g[key].shape), g[pre + "_d_" + method], **TOL, name="d")
        assert_close(h.reshape(g[key].shape), np.real(g[pre + "_h_" + method]), **TOL, name="h")


@pytest.mark.parametrize("mi,method", list(enumerate(O.METHODS)))
def test_regularize_inverse_fused(hm, mi, method):
    """The fused d = regularize(1/b) (one division on the fast path) against the reference sequence,
    including |b| far outside the fast range, zeros, Inf and NaN."""
    rng = np.random.default_rng(7)
    b = (rng.standard_normal(400) + 1j * rng.standard_normal(400)) * 10.0 ** rng.uniform(-30, 30, 400)
    b = np.concatenate([b, [1e-200 + 1e-210j, 1e200j, 3e-160, 0j, np.inf, complex(np.nan, 1.0), 1e-310 + 0j]])
    with np.errstate(all="ignore"):
        ref_d, ref_h = O.regularize(1 / b, 100.0, method)
    d, h = np.empty_like(b), np.empty(b.shape)
    hm.hm_regularize_inverse(b.size, P(np.ascontiguousarray(b)), 100.0, mi, P(d), P(h))
    assert_close(h, np.real(ref_h), rtol=1e-12, floor=0.0, name="h")
    assert_close(d, ref_d, rtol=1e-12, floor=0.0, name="d")
