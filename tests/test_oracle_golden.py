"""Pin the oracle (oracle/sfa_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import sfa_oracle as O
from conftest import assert_close

EXACT = dict(rtol=0.0, floor=0.0)
TIGHT = dict(rtol=1e-14, floor=1e-16)


def test_sht_matrix(golden):
    g = golden["angular"]
    assert_close(O.sht_matrix(4, g["ea4_azi"], g["ea4_colat"], g["ea4_w"]), g["ea4_Y"], **EXACT, name="ea4")
    assert_close(O.sht_matrix(4, g["g4_azi"], g["g4_colat"]), g["g4_Y"], **EXACT, name="g4")
    assert_close(O.sht_matrix(8, g["r8_azi"], g["r8_colat"]), g["r8_Y"], **EXACT, name="r8")
    assert_close(O.sht_matrix(8, g["eq8_azi"], np.pi / 2), g["eq8_Y"], **EXACT, name="eq8")
    assert_close(O.sht_matrix(32, g["r32_azi"], g["r32_colat"]), g["r32_Y"], **EXACT, name="r32")
    assert_close(O.sht_matrix(0, 0.3, 0.4), g["s0_Y"], **EXACT, name="s0")
    with pytest.raises(ValueError):
        O.sht_matrix(1, 0.3, [0.1, 0.2])
    with pytest.raises(ValueError):
        O.sht_matrix(1, np.zeros((2, 2)), np.zeros((2, 2)))


def test_cht_matrix(golden):
    g = golden["angular"]
    assert_close(O.cht_matrix(16, g["c16_pol"], g["c16_w"]), g["c16_Psi"], **EXACT, name="c16")
    assert_close(O.cht_matrix(5, g["c5_pol"]), g["c5_Psi"], **EXACT, name="c5")
    assert_close(O.cht_matrix(0, 1.25), g["c0_Psi"], **EXACT, name="c0")
    assert list(O.ch_order(3)) == [0, 1, 2, 3, -3, -2, -1]


@pytest.mark.parametrize("setup", O.SETUPS)
def test_radial_weights(golden, setup):
    g = golden["radial"]
    assert_close(O.weights(8, g["kr8"], setup), g["w8_" + setup], **EXACT, name="w8")
    assert_close(O.weights(32, g["kr32"], setup), g["w32_" + setup], **EXACT, name="w32")
    assert_close(O.circ_radial_weights(16, g["kr8"], setup), g["cw16_" + setup], **EXACT, name="cw16")
    assert_close(O.circ_radial_weights(32, g["kr32"], setup), g["cw32_" + setup], **EXACT, name="cw32")


@pytest.mark.parametrize("setup", O.SETUPS)
def test_radial_filters(golden, setup):
    g = golden["radial"]
    k = g["k40"]
    assert_close(O.spherical_pw(4, k, 0.042, setup), g["pw4_" + setup], **EXACT, name="pw")
    assert_close(O.spherical_pw(4, k, 0.042, setup, replace_zeros=False), g["pw4nr_" + setup], **EXACT, name="pwnr")
    assert_close(O.spherical_ps(4, k, 0.042, 1.5, setup), g["ps4_" + setup], **EXACT, name="ps")
    assert_close(O.circular_pw(16, k, 0.1, setup), g["cpw16_" + setup], **EXACT, name="cpw")
    assert_close(O.circular_ls(6, k[1:], 0.1, 2.0, setup), g["cls6_" + setup], **EXACT, name="cls")


def test_shapes_and_zero_replacement(golden):
    g = golden["radial"]
    assert_close(O.spherical_pw(3, g["ku"], 1.0, "open"), g["pw3_unsorted"], **EXACT, name="unsorted")
    assert_close(O.circular_pw(3, g["ku"], 1.0, "card"), g["cpw3_unsorted"], **EXACT, name="cunsorted")
    assert O.weights(3, 0.5, "open").shape == (4,)
    assert_close(O.weights(3, 0.5, "open"), g["w_scalar"], **EXACT)
    assert_close(O.weights(0, [0.5, 0.6], "open"), g["w_n0"], **EXACT)
    assert_close(O.spherical_pw(5, 2.0, 0.5, "rigid"), g["pw_scalar"], **EXACT)
    assert_close(O.replace_zeros_of_radial_function(g["rz_A"], g["rz_kr"]), g["rz_out"], **EXACT)
    with pytest.raises(ValueError):
        O.replace_zeros_of_radial_function(np.zeros((3, 2)), np.arange(3.0))
    with pytest.raises(ValueError):
        O.replace_zeros_of_radial_function(np.ones((3, 2)), np.arange(4.0))
    with pytest.raises(ValueError):
        O.weights(2, 1.0, "closed")
    # A7: documented value
    pw = O.spherical_pw(2, np.array([0.0, 1.0]), 1, "open")
    assert abs(pw[0, 0] - 4 * np.pi) < 1e-14 and abs(pw[0, 1] - 3.7845972j) < 1e-6


@pytest.mark.parametrize("method", O.METHODS)
def test_regularize(golden, method):
    g = golden["radial"]
    d, h = O.regularize(g["reg_dn"], 100.0, method)
    assert_close(d, g["reg_d_" + method], **EXACT, name="d")
    assert_close(h, g["reg_h_" + method], **EXACT, name="h")
    assert h.dtype == g["reg_h_" + method].dtype
    if method in ("softclip", "Tikh", "wng", "hardclip", "discard"):
        d, h = O.regularize(g["reg_big"], 10.0, method)
        assert_close(d, g["regbig_d_" + method], **EXACT, name="dbig")
        assert_close(h, g["regbig_h_" + method], **EXACT, name="hbig")
    if method == "Tikh":
        d, h = O.regularize(g["reg_dn"], 1.0, "Tikh")
        assert np.isnan(h).all() and np.isnan(d).all()
    with pytest.raises(ValueError):
        O.regularize(g["reg_dn"], 1.0, "tikhonov")


def test_mode_matrices(golden):
    g = golden["radial"]
    v = g["rep_in"]
    assert_close(O.repeat_n_m(v), g["rep_out"], **EXACT)
    assert_close(O.repeat_n_m(v[0]), g["rep_out_1d"], **EXACT)
    assert_close(O.diagonal_mode_mat(v[:, :3]), g["diag_out"], **EXACT)
    assert_close(O.diagonal_mode_mat(v[0, :3]), g["diag_out_1"], **EXACT)
    assert_close(O.circ_diagonal_mode_mat(v[:, :3]), g["cdiag_out"], **EXACT)


def test_pwd_chain(golden):
    g = golden["pwd"]
    N = 4
    Y_p = O.sht_matrix(N, g["sph_azi"], g["sph_colat"], g["sph_w"])
    Y_q = O.sht_matrix(N, g["sph_azi_l"], np.pi / 2)
    bn = O.spherical_pw(N, g["sph_k"], 0.042, "rigid")
    dn, _ = O.regularize(1 / bn, 100, "softclip")
    assert_close(dn, g["sph_dn"], **EXACT)
    q = O.pwd_literal(Y_q, O.diagonal_mode_mat(dn), Y_p, g["sph_X"])
    assert_close(q, g["sph_q"], **TIGHT, name="literal")
    q2 = O.pwd_factored(Y_q, O.repeat_n_m(dn), Y_p, g["sph_X"])
    assert np.max(np.abs(q2 - g["sph_q"])) <= 1e-12 * np.max(np.abs(g["sph_q"]))
    N = 6
    Psi_p = O.cht_matrix(N, g["circ_pol"], g["circ_w"])
    Psi_q = O.cht_matrix(N, g["circ_pol_l"])
    Bn = O.circular_pw(N, g["sph_k"], 0.1, "open")
    Dn, _ = O.regularize(1 / Bn, 3000, "softclip")
    assert_close(Dn, g["circ_dn"], **EXACT)
    q = O.pwd_factored(Psi_q, Dn, Psi_p, g["circ_X"])
    assert np.max(np.abs(q - g["circ_q"])) <= 1e-12 * np.max(np.abs(g["circ_q"]))


def _dot_sph(v, u):
    return (np.cos(v[0]) * np.sin(v[1]) * np.cos(u[0]) * np.sin(u[1]) +
            np.sin(v[0]) * np.sin(v[1]) * np.sin(u[0]) * np.sin(u[1]) +
            np.cos(v[1]) * np.cos(u[1]))


def test_example_fingerprints(golden):
    """peak index at bin 50 and sum|q| of the four example scripts."""
    g = golden["examples"]
    expect = {"open_circular": (111, 32842.6924255), "rigid_circular": (90, 30455.3171289),
              "open_spherical": (45, 9718.40942342), "rigid_spherical": (45, 10197.2429958)}
    for name, (peak, total) in expect.items():
        q = g[name + "_q"]
        assert int(np.argmax(np.abs(q[50]))) == peak
        assert abs(np.sum(np.abs(q)) - total) < 1e-6 * total
    # re-run the example pipelines through the oracle
    k = g["open_spherical_k"]
    azi, colat, w = O.grid_gauss(20)
    azi_l = np.linspace(0, 2 * np.pi, 91, endpoint=False)
    Y_p = O.sht_matrix(20, azi, colat, w)
    Y_q = O.sht_matrix(20, azi_l, np.pi / 2)
    p = np.exp(1j * k[:, None] * 1 * _dot_sph((azi, colat), (np.pi, np.pi / 2)))
    dn, _ = O.regularize(1 / O.spherical_pw(20, k, 1, "open"), 100, "softclip")
    q = np.squeeze(O.pwd_factored(Y_q, O.repeat_n_m(dn), Y_p, p))
    assert np.max(np.abs(q - g["open_spherical_q"])) <= 1e-11 * np.max(np.abs(q))
    dn, _ = O.regularize(1 / O.spherical_pw(20, k, 1, "rigid"), 100, "softclip")
    q = np.squeeze(O.pwd_factored(Y_q, O.repeat_n_m(dn), Y_p, g["rigid_spherical_p"]))
    assert np.max(np.abs(q - g["rigid_spherical_q"])) <= 1e-11 * np.max(np.abs(q))
    pol = np.linspace(0, 2 * np.pi, 61, endpoint=False)
    pol_l = np.linspace(0, 2 * np.pi, 180, endpoint=False)
    Psi_p = O.cht_matrix(30, pol, np.full(61, 1 / 61))
    Psi_q = O.cht_matrix(30, pol_l)
    for name, a0 in (("open", 3000), ("rigid", 100)):
        Dn, _ = O.regularize(1 / O.circular_pw(30, k, 1, name), a0, "softclip")
        q = np.squeeze(O.pwd_factored(Psi_q, Dn, Psi_p, g[name + "_circular_p"]))
        assert np.max(np.abs(q - g[name + "_circular_q"])) <= 1e-11 * np.max(np.abs(q))


def test_identities_mpmath():
    """Independent truth: a few mpmath values and analytic identities."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    azi, colat, w = O.grid_gauss(3)
    Y = O.sht_matrix(3, azi, colat)
    G = Y.conj().T @ (w[:, None] * Y)
    assert np.max(np.abs(G - np.eye(16))) < 1e-13
    Y = O.sht_matrix(5, [0.7], [1.1])
    n, m = O.sh_index(5)
    for i in (0, 3, 10, 24, 35):
        ref = complex(mp.spherharm(int(n[i]), int(m[i]), 1.1, 0.7))
        assert abs(Y[0, i] - ref) <= 1e-13 * abs(ref)
    # rigid closed form b_n = -i / ((kr)^2 h_n^(2)'(kr))
    x = 2.3
    b = O.weights(4, x, "rigid")
    for nn in range(5):
        h = lambda t: mp.sqrt(mp.pi / (2 * t)) * (mp.besselj(nn + 0.5, t) - 1j * mp.bessely(nn + 0.5, t))
        hd = mp.diff(h, x)
        ref = complex(-1j / (x ** 2 * hd))
        assert abs(b[nn] - ref) <= 1e-12 * abs(ref)


def test_extras(golden):
    """SURVEY 8f "next" rows against the unmodified reference."""
    g = golden["extras"]
    assert_close(O.Legendre_matrix(12, g["leg_ct"]), g["leg12"], **EXACT)
    assert_close(O.matdiagmul(g["mdm_A"], g["mdm_b"]), g["mdm_C"], **EXACT)
    assert_close(O.matdiagmul(g["mdm_A"], g["mdm_b"][0]), g["mdm_C1"], **EXACT)
    assert_close(O.norm_of_columns(g["mdm_A"]), g["noc2"], **TIGHT)
    assert_close(O.norm_of_columns(g["mdm_A"], 1), g["noc1"], **TIGHT)
    assert abs(O.coherence_of_columns(g["mdm_A"]) - float(g["coh"])) < 1e-14
    assert abs(O.coherence_of_columns(g["coh_Y_in"]) - float(g["coh_Y"])) < 1e-14


def test_oracle_irfft_shift_matches_examples(golden):
    """The examples' last step, q_t = fftshift(irfft(q, axis=0), axes=0): the oracle restatement applied to the
    golden q of the unmodified scripts reproduces their q_t bit-exactly (same NumPy)."""
    import os
    gt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples_time.npz"))
    for name in ("open_circular", "rigid_circular", "open_spherical", "rigid_spherical"):
        out = O.irfft_shift(golden["examples"][name + "_q"])
        assert out.shape == gt[name + "_q_t"].shape == (198, golden["examples"][name + "_q"].shape[1])
        assert np.array_equal(out, gt[name + "_q_t"])
