#!/usr/bin/env python
"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``  -> ``tests/golden/*.npz``.

The reference is imported from ``/root/reference`` with the one shim SURVEY.md
section 8c documents (``scipy.special.sph_harm`` is gone from SciPy >= 1.17;
it forwarded to ``sph_harm_y(n, m, colat, azi)``) and a ``matplotlib`` stub so
that the four example scripts can be executed as they are.
Versions used are stored in ``meta.npz``.
"""
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import scipy
import scipy.special as sp

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

sp.sph_harm = lambda m, n, theta, phi: sp.sph_harm_y(n, m, phi, theta)
sys.path.insert(0, REF)
import micarray  # noqa: E402
from micarray.modal import angular, radial  # noqa: E402


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def golden_angular():
    out = {}
    azi, colat, w = angular.grid_equal_angle(4)
    out["ea4_azi"], out["ea4_colat"], out["ea4_w"] = azi, colat, w
    out["ea4_Y"] = angular.sht_matrix(4, azi, colat, w)
    azi, colat, w = angular.grid_gauss(4)
    out["g4_azi"], out["g4_colat"], out["g4_w"] = azi, colat, w
    out["g4_Y"] = angular.sht_matrix(4, azi, colat)
    rng = np.random.default_rng(1)
    azi = rng.uniform(0, 2 * np.pi, 37)
    colat = np.arccos(rng.uniform(-1, 1, 37))
    colat[:6] = [0.0, np.pi, np.pi / 2, 1e-9, np.pi - 1e-9, 1e-3]
    out["r8_azi"], out["r8_colat"] = azi, colat
    out["r8_Y"] = angular.sht_matrix(8, azi, colat)
    out["eq8_azi"] = np.linspace(0, 2 * np.pi, 24, endpoint=False)
    out["eq8_Y"] = angular.sht_matrix(8, out["eq8_azi"], np.pi / 2)
    azi = rng.uniform(-7, 7, 9)
    colat = np.arccos(rng.uniform(-1, 1, 9))
    out["r32_azi"], out["r32_colat"] = azi, colat
    out["r32_Y"] = angular.sht_matrix(32, azi, colat)
    pol, w = angular.grid_equal_polar_angle(16)
    out["c16_pol"], out["c16_w"] = pol, w
    out["c16_Psi"] = angular.cht_matrix(16, pol, w)
    out["c5_pol"] = rng.uniform(-10, 10, 11)
    out["c5_Psi"] = angular.cht_matrix(5, out["c5_pol"])
    out["c0_Psi"] = angular.cht_matrix(0, 1.25)
    out["s0_Y"] = angular.sht_matrix(0, 0.3, 0.4)
    save("angular", **out)


def golden_radial():
    out = {}
    kr8 = np.array([0.0, 1e-6, 1e-3, 0.01, 0.1, 0.5, 1.0, 2.0, 4.493409457909064,
                    5.0, 8.9, 10.0, 20.0, 35.5, 50.0])
    kr32 = np.geomspace(0.01, 60.0, 24)
    out["kr8"], out["kr32"] = kr8, kr32
    for setup in ("open", "card", "rigid"):
        out["w8_" + setup] = radial.weights(8, kr8, setup)
        out["w32_" + setup] = radial.weights(32, kr32, setup)
        out["cw16_" + setup] = radial.circ_radial_weights(16, kr8, setup)
        out["cw32_" + setup] = radial.circ_radial_weights(32, kr32, setup)
    k = 2 * np.pi * np.linspace(0, 24000, 40, endpoint=False) / 343
    out["k40"] = k
    with np.errstate(all="ignore"):
        for setup in ("open", "card", "rigid"):
            out["pw4_" + setup] = radial.spherical_pw(4, k, 0.042, setup)
            out["pw4nr_" + setup] = radial.spherical_pw(4, k, 0.042, setup, replace_zeros=False)
            out["ps4_" + setup] = radial.spherical_ps(4, k, 0.042, 1.5, setup)
            out["cpw16_" + setup] = radial.circular_pw(16, k, 0.1, setup)
            out["cls6_" + setup] = radial.circular_ls(6, k[1:], 0.1, 2.0, setup)
    # unsorted + duplicated wavenumbers, zero in the middle
    ku = np.array([3.0, 0.0, 1.0, 3.0, 0.5, 0.0, 7.0])
    out["ku"] = ku
    out["pw3_unsorted"] = radial.spherical_pw(3, ku, 1.0, "open")
    out["cpw3_unsorted"] = radial.circular_pw(3, ku, 1.0, "card")
    out["w_scalar"] = radial.weights(3, 0.5, "open")
    out["w_n0"] = radial.weights(0, [0.5, 0.6], "open")
    out["pw_scalar"] = radial.spherical_pw(5, 2.0, 0.5, "rigid")
    # synthetic replace_zeros case with ties and runs of zeros
    A = np.arange(1, 25, dtype=float).reshape(6, 4) * (1 + 0.5j)
    A[0, 1] = A[1, 1] = 0
    A[2, 2] = 0
    A[3:, 3] = 0
    A[4, 0] = 1e-301
    krz = np.array([0.0, 1.0, 2.0, 3.0, 4.0, 5.0])
    out["rz_A"], out["rz_kr"] = A, krz
    out["rz_out"] = radial.replace_zeros_of_radial_function(A.copy(), krz)
    # regularisers on a wide-dynamic-range inverse
    bn = radial.spherical_pw(8, np.geomspace(0.01, 50, 32) / 0.1, 0.1, "card")
    dn = 1 / bn
    out["reg_dn"] = dn
    with np.errstate(all="ignore"):
        for method in ("none", "discard", "hardclip", "softclip", "Tikh", "wng"):
            d, h = radial.regularize(dn, 100.0, method)
            out["reg_d_" + method], out["reg_h_" + method] = d, h
        d, h = radial.regularize(dn, 1.0, "Tikh")       # a0 < 2 -> NaN
        out["reg_d_Tikh_nan"], out["reg_h_Tikh_nan"] = d, h
        big = np.array([1e200 + 0j, 1e-200j, 3 - 4j, np.inf, 0j])
        out["reg_big"] = big
        for method in ("softclip", "Tikh", "wng", "hardclip", "discard"):
            d, h = radial.regularize(big, 10.0, method)
            out["regbig_d_" + method], out["regbig_h_" + method] = d, h
    v = np.arange(12, dtype=float).reshape(3, 4) + 1j
    out["rep_in"] = v
    out["rep_out"] = radial.repeat_n_m(v)
    out["rep_out_1d"] = radial.repeat_n_m(v[0])
    out["diag_out"] = radial.diagonal_mode_mat(v[:, :3])
    out["diag_out_1"] = radial.diagonal_mode_mat(v[0, :3])
    out["cdiag_out"] = radial.circ_diagonal_mode_mat(v[:, :3])
    save("radial", **out)


def golden_pwd():
    """The example idiom (rigid_spherical_array.py:18-39) at a small size."""
    out = {}
    N, F, T = 4, 12, 3
    rng = np.random.default_rng(2)
    azi, colat, w = angular.grid_gauss(N)
    k = 2 * np.pi * np.linspace(0, 8000, F, endpoint=False) / 343
    azi_l = np.linspace(0, 2 * np.pi, 20, endpoint=False)
    bn = radial.spherical_pw(N, k, 0.042, setup="rigid")
    dn, _ = radial.regularize(1 / bn, 100, "softclip")
    D = radial.diagonal_mode_mat(dn)
    Y_p = angular.sht_matrix(N, azi, colat, w)
    Y_q = angular.sht_matrix(N, azi_l, np.pi / 2)
    X = (rng.standard_normal((F, len(azi), T)) + 1j * rng.standard_normal((F, len(azi), T)))
    A = np.matmul(np.matmul(Y_q, D), np.conj(Y_p.T))
    q = np.matmul(A, X)
    out.update(sph_k=k, sph_azi=azi, sph_colat=colat, sph_w=w, sph_azi_l=azi_l, sph_X=X,
               sph_dn=dn, sph_A=A, sph_q=q)
    # circular (open_circular_array.py:33-39)
    N = 6
    pol, w = angular.grid_equal_polar_angle(N)
    pol_l = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    Bn = radial.circular_pw(N, k, 0.1, setup="open")
    Dn, _ = radial.regularize(1 / Bn, 3000, "softclip")
    D = radial.circ_diagonal_mode_mat(Dn)
    Psi_p = angular.cht_matrix(N, pol, w)
    Psi_q = angular.cht_matrix(N, pol_l)
    X = (rng.standard_normal((F, len(pol), T)) + 1j * rng.standard_normal((F, len(pol), T)))
    q = np.matmul(np.matmul(np.matmul(Psi_q, D), np.conj(Psi_p.T)), X)
    out.update(circ_pol=pol, circ_w=w, circ_pol_l=pol_l, circ_X=X, circ_dn=Dn, circ_q=q)
    save("pwd", **out)


def golden_examples():
    """Execute the four reference example scripts unmodified; keep q and p."""
    plt = types.ModuleType("matplotlib.pyplot")
    for fn in ("figure", "pcolormesh", "colorbar", "xlabel", "ylabel", "title", "savefig",
               "plot", "show", "legend", "grid", "xlim", "ylim", "subplot", "imshow"):
        setattr(plt, fn, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    out = {}
    out_t = {}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        for name, qvar in (("open_circular", "q_pwd"), ("rigid_circular", "q_pwd"),
                           ("open_spherical", "q_pwd"), ("rigid_spherical", "q_mb")):
            path = os.path.join(REF, "doc", "examples", "modal_beamforming_%s_array.py" % name)
            with np.errstate(all="ignore"):
                g = runpy.run_path(path)
            out[name + "_q"] = np.asarray(g[qvar])
            # the examples' last step: q_t = fftshift(irfft(q, axis=0), axes=0)  (e.g. ..._rigid_spherical_array.py:40)
            out_t[name + "_q_t"] = np.asarray(g[qvar + "_t"])
            if name != "open_spherical":      # closed form; rebuilt in the test (1.4 MB saved)
                out[name + "_p"] = np.asarray(g["p"])
            out[name + "_k"] = np.asarray(g["k"], dtype=float)
    finally:
        os.chdir(cwd)
    save("examples", **out)
    save("examples_time", **out_t)


def golden_extras():
    """SURVEY 8f "next" rows: Legendre_matrix and the util helpers."""
    from micarray import util
    out = {}
    rng = np.random.default_rng(9)
    ct = np.cos(np.linspace(0, np.pi, 41))
    out["leg_ct"] = ct
    out["leg12"] = angular.Legendre_matrix(12, ct)
    A = rng.standard_normal((7, 5)) + 1j * rng.standard_normal((7, 5))
    b = rng.standard_normal((3, 5)) + 1j * rng.standard_normal((3, 5))
    out["mdm_A"], out["mdm_b"] = A, b
    out["mdm_C"] = util.matdiagmul(A, b)
    out["mdm_C1"] = util.matdiagmul(A, b[0])
    out["noc2"] = util.norm_of_columns(A)
    out["noc1"] = util.norm_of_columns(A, p=1)
    out["coh"] = np.asarray(util.coherence_of_columns(A))
    Y = angular.sht_matrix(3, *angular.grid_gauss(3)[:2])
    out["coh_Y"] = np.asarray(util.coherence_of_columns(Y))
    out["coh_Y_in"] = Y
    save("extras", **out)


if __name__ == "__main__":
    golden_extras()
    golden_angular()
    golden_radial()
    golden_pwd()
    golden_examples()
    save("meta", scipy=np.array(scipy.__version__), numpy=np.array(np.__version__),
         reference=np.array("spatialaudio/sfa-numpy micarray " + micarray.__version__))
