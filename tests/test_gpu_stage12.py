"""GPU parity of stages 1-2 (SH/CH matrices, radial filters, regulariser) against the oracle
and the golden vectors of the unmodified reference.  Everything goes through the C ABI
(libsfa_b200.so via sfa_numpy_b200._lib).  Tolerance (north_star): relative error <= 1e-10,
with the absolute floor 1e-15 * column max of SURVEY 7 for entries near zeros of the functions."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, floor=1e-15)


@pytest.fixture(scope="module")
def mic():
    import sfa_numpy_b200 as micarray
    return micarray


def cpu(t):
    return t.detach().cpu().numpy()


def test_sht_golden(mic, golden):
    g = golden["angular"]
    A = mic.modal.angular
    assert_close(cpu(A.sht_matrix(4, g["ea4_azi"], g["ea4_colat"], g["ea4_w"])), g["ea4_Y"], **TOL, name="ea4")
    assert_close(cpu(A.sht_matrix(4, g["g4_azi"], g["g4_colat"])), g["g4_Y"], **TOL, name="g4")
    assert_close(cpu(A.sht_matrix(8, g["r8_azi"], g["r8_colat"])), g["r8_Y"], **TOL, name="r8")
    assert_close(cpu(A.sht_matrix(8, g["eq8_azi"], np.pi / 2)), g["eq8_Y"], **TOL, name="eq8")
    assert_close(cpu(A.sht_matrix(32, g["r32_azi"], g["r32_colat"])), g["r32_Y"], **TOL, name="r32")
    assert_close(cpu(A.sht_matrix(0, 0.3, 0.4)), g["s0_Y"], **TOL, name="s0")
    Y = A.sht_matrix(3, [0.1, 0.2], [0.3, 0.4])
    assert Y.shape == (2, 16) and str(Y.dtype) == "torch.complex128" and Y.is_cuda
    with pytest.raises(ValueError):
        A.sht_matrix(1, 0.3, [0.1, 0.2])
    with pytest.raises(ValueError):
        A.sht_matrix(1, np.zeros((2, 2)), np.zeros((2, 2)))


@pytest.mark.parametrize("N,P", [(1, 33), (2, 100), (5, 700), (8, 2562), (16, 2562), (17, 311), (32, 1089), (40, 64)])
def test_sht_oracle(mic, N, P):
    azi, colat, w = O.fibonacci_grid(P)
    azi = azi - np.pi            # negative azimuths too
    Y = cpu(mic.modal.angular.sht_matrix(N, azi, colat, w))
    assert_close(Y, O.sht_matrix(N, azi, colat, w), **TOL, name="fib N=%d" % N)
    # orthonormality on the quadrature grid (identity, independent of the oracle)
    if N <= 8:
        az, co, ww = O.grid_gauss(N)
        Yg = cpu(mic.modal.angular.sht_matrix(N, az, co))
        G = Yg.conj().T @ (ww[:, None] * Yg)
        assert np.max(np.abs(G - np.eye((N + 1) ** 2))) < 1e-12


def test_cht(mic, golden):
    g = golden["angular"]
    A = mic.modal.angular
    assert_close(cpu(A.cht_matrix(16, g["c16_pol"], g["c16_w"])), g["c16_Psi"], **TOL, name="c16")
    assert_close(cpu(A.cht_matrix(5, g["c5_pol"])), g["c5_Psi"], **TOL, name="c5")
    assert_close(cpu(A.cht_matrix(0, 1.25)), g["c0_Psi"], **TOL, name="c0")
    pol = np.linspace(-20, 20, 1000)
    assert_close(cpu(A.cht_matrix(30, pol)), O.cht_matrix(30, pol), **TOL, name="c30")
    # many rows: the shared-memory row-block kernel (recurrence in n), ragged last block, weights
    rng = np.random.default_rng(7)
    for N, Q in ((16, 5000), (1, 4097), (40, 4100), (100, 4096)):
        pol = rng.uniform(-4 * np.pi, 4 * np.pi, Q)
        w = rng.uniform(0.1, 2.0, Q)
        assert_close(cpu(A.cht_matrix(N, pol, w)), O.cht_matrix(N, pol, w), **TOL, name="rows N=%d" % N)
    assert_close(cpu(A.cht_matrix(16, pol[:4096])), O.cht_matrix(16, pol[:4096]), **TOL, name="rows no weights")


@pytest.mark.parametrize("setup", O.SETUPS)
def test_radial_golden(mic, golden, setup):
    g = golden["radial"]
    R = mic.modal.radial
    assert_close(cpu(R.weights(8, g["kr8"], setup)), g["w8_" + setup], **TOL, name="w8")
    assert_close(cpu(R.weights(32, g["kr32"], setup)), g["w32_" + setup], **TOL, name="w32")
    assert_close(cpu(R.circ_radial_weights(16, g["kr8"], setup)), g["cw16_" + setup], **TOL, name="cw16")
    assert_close(cpu(R.circ_radial_weights(32, g["kr32"], setup)), g["cw32_" + setup], **TOL, name="cw32")
    k = g["k40"]
    assert_close(cpu(R.spherical_pw(4, k, 0.042, setup)), g["pw4_" + setup], **TOL, name="pw")
    assert_close(cpu(R.spherical_pw(4, k, 0.042, setup, replace_zeros=False)), g["pw4nr_" + setup], **TOL, name="pwnr")
    assert_close(cpu(R.spherical_ps(4, k, 0.042, 1.5, setup)), g["ps4_" + setup], **TOL, name="ps")
    assert_close(cpu(R.circular_pw(16, k, 0.1, setup)), g["cpw16_" + setup], **TOL, name="cpw")
    assert_close(cpu(R.circular_ls(6, k[1:], 0.1, 2.0, setup)), g["cls6_" + setup], **TOL, name="cls")


@pytest.mark.parametrize("setup", O.SETUPS)
@pytest.mark.parametrize("N,F", [(8, 1024), (16, 1024), (32, 4096), (4, 512)])
def test_radial_oracle_sweeps(mic, setup, N, F):
    R = mic.modal.radial
    kr = np.geomspace(0.01, 50, F)                       # cfg 4 sweep
    assert_close(cpu(R.weights(N, kr, setup)), O.weights(N, kr, setup), **TOL, name="sph")
    assert_close(cpu(R.circ_radial_weights(N, kr, setup)), O.circ_radial_weights(N, kr, setup), **TOL, name="circ")
    k = O.wavenumbers(F)                                 # includes k = 0 -> zero replacement
    with np.errstate(all="ignore"):
        assert_close(cpu(R.spherical_pw(N, k, 0.042, setup)), O.spherical_pw(N, k, 0.042, setup), **TOL, name="pw")
        assert_close(cpu(R.circular_pw(N, k, 0.1, setup)), O.circular_pw(N, k, 0.1, setup), **TOL, name="cpw")


def test_radial_shapes_errors(mic, golden):
    g = golden["radial"]
    R = mic.modal.radial
    assert tuple(R.weights(3, 0.5, "open").shape) == (4,)
    assert tuple(R.weights(0, [0.5, 0.6], "open").shape) == (2,)
    assert_close(cpu(R.weights(3, 0.5, "open")), g["w_scalar"], **TOL)
    assert_close(cpu(R.spherical_pw(5, 2.0, 0.5, "rigid")), g["pw_scalar"], **TOL)
    assert_close(cpu(R.spherical_pw(3, g["ku"], 1.0, "open")), g["pw3_unsorted"], **TOL, name="unsorted")
    assert_close(cpu(R.circular_pw(3, g["ku"], 1.0, "card")), g["cpw3_unsorted"], **TOL, name="cunsorted")
    assert_close(cpu(R.replace_zeros_of_radial_function(g["rz_A"], g["rz_kr"])), g["rz_out"], rtol=0, floor=0)
    with pytest.raises(ValueError):
        R.weights(2, 1.0, "closed")
    with pytest.raises(ValueError):
        R.replace_zeros_of_radial_function(np.zeros((3, 2)), np.arange(3.0))
    with pytest.raises(ValueError):
        R.replace_zeros_of_radial_function(np.ones((3, 2)), np.arange(4.0))
    with pytest.raises(ValueError):
        R.spherical_pw(2, np.zeros(3), 1.0, "open")       # every column n >= 1 is all zero
    with pytest.raises(TypeError):
        R.spherical_pw(2, [0.0, 1.0], 0.5, "open")        # list * float, as in the reference


@pytest.mark.parametrize("method", O.METHODS)
def test_regularize(mic, golden, method):
    g = golden["radial"]
    R = mic.modal.radial
    d, h = R.regularize(g["reg_dn"], 100.0, method)
    assert_close(cpu(d), g["reg_d_" + method], **TOL, name="d")
    assert_close(cpu(h), g["reg_h_" + method], **TOL, name="h")
    assert cpu(h).dtype == g["reg_h_" + method].dtype
    if method != "none":
        d, h = R.regularize(g["reg_big"], 10.0, method)
        assert_close(cpu(d), g["regbig_d_" + method], **TOL, name="dbig")
        assert_close(cpu(h), g["regbig_h_" + method], **TOL, name="hbig")
    if method == "Tikh":
        d, h = R.regularize(g["reg_dn"], 1.0, "Tikh")
        assert np.isnan(cpu(h)).all()
    with pytest.raises(ValueError):
        R.regularize(g["reg_dn"], 1.0, "tikhonov")


@pytest.mark.parametrize("method", O.METHODS)
def test_fused_radial_filters(mic, method):
    """bn, dn, hn from ONE kernel == reference sequence spherical_pw -> regularize(1/bn), for all six
    regularisers (radial.py:240-266): both branches of the fused inverse (one-division closed form for
    none/Tikh/wng, Smith division for the clipping methods) and both hn store types (complex for
    none/discard/hardclip, real otherwise)."""
    k = O.wavenumbers(1024)
    bn, dn, hn = mic.modal.radial.radial_filters(8, k, 0.042, "rigid", 100.0, method)
    ref_b = O.spherical_pw(8, k, 0.042, "rigid")
    with np.errstate(all="ignore"):
        ref_d, ref_h = O.regularize(1 / ref_b, 100.0, method)
    assert_close(cpu(bn), ref_b, **TOL, name="bn")
    assert_close(cpu(dn), ref_d, **TOL, name="dn")
    assert_close(cpu(hn), ref_h, **TOL, name="hn")
    assert cpu(hn).dtype == ref_h.dtype
    bn, dn, hn = mic.modal.radial.radial_filters(16, k, 0.1, "open", 3000.0, method, kind="circular_pw")
    ref_b = O.circular_pw(16, k, 0.1, "open")
    with np.errstate(all="ignore"):
        ref_d, ref_h = O.regularize(1 / ref_b, 3000.0, method)
    assert_close(cpu(dn), ref_d, **TOL, name="cdn")
    assert_close(cpu(hn), ref_h, **TOL, name="chn")
    assert cpu(hn).dtype == ref_h.dtype


@pytest.mark.parametrize("method", O.METHODS)
def test_cfg4_radial_all_methods(mic, method):
    """cardioid sphere N=16, kr = geomspace(0.01, 50, 1024) (min |b_n| ~ 2.5e-48,
    1/b ~ 4e47), every regulariser, through the fused kernel."""
    r = 0.1
    k = np.geomspace(0.01, 50, 1024) / r
    bn, dn, hn = mic.modal.radial.radial_filters(16, k, r, "card", 100.0, method)
    ref_b = O.spherical_pw(16, k, r, "card")
    with np.errstate(all="ignore"):
        ref_d, ref_h = O.regularize(1 / ref_b, 100.0, method)
    assert_close(cpu(bn), ref_b, **TOL, name="bn")
    assert_close(cpu(dn), ref_d, **TOL, name="dn")
    assert_close(cpu(hn), ref_h, **TOL, name="hn")
    assert cpu(hn).dtype == ref_h.dtype


def test_mode_matrices(mic, golden):
    g = golden["radial"]
    R = mic.modal.radial
    v = g["rep_in"]
    assert_close(cpu(R.repeat_n_m(v)), g["rep_out"], rtol=0, floor=0)
    assert_close(cpu(R.repeat_n_m(v[0])), g["rep_out_1d"], rtol=0, floor=0)
    assert_close(cpu(R.diagonal_mode_mat(v[:, :3])), g["diag_out"], rtol=0, floor=0)
    assert_close(cpu(R.diagonal_mode_mat(v[0, :3])), g["diag_out_1"], rtol=0, floor=0)
    assert_close(cpu(R.circ_diagonal_mode_mat(v[:, :3])), g["cdiag_out"], rtol=0, floor=0)


def test_empty_and_degenerate_inputs(mic):
    """Edge cases: no points / no bins / order 0 -- same shapes as the reference (oracle)."""
    A, R = mic.modal.angular, mic.modal.radial
    Y = A.sht_matrix(3, [], [])
    assert tuple(Y.shape) == O.sht_matrix(3, [], []).shape == (0, 16)
    assert tuple(A.cht_matrix(2, []).shape) == O.cht_matrix(2, []).shape == (0, 5)
    assert tuple(R.weights(4, [], "rigid").shape) == O.weights(4, [], "rigid").shape
    assert_close(cpu(A.sht_matrix(0, [0.1, 0.2, 0.3], [1.0, 1.1, 1.2])), O.sht_matrix(0, [0.1, 0.2, 0.3], [1.0, 1.1, 1.2]), **TOL)
    assert_close(cpu(R.weights(0, [0.0, 0.5, 7.0], "rigid")), O.weights(0, [0.0, 0.5, 7.0], "rigid"), **TOL)
    assert_close(cpu(R.circular_pw(0, np.array([0.0, 0.5, 7.0]), 1.0, "card")), O.circular_pw(0, np.array([0.0, 0.5, 7.0]), 1.0, "card"), **TOL)
    # one very large order (table in shared memory) and one very large argument
    az, co, w = O.fibonacci_grid(50)
    assert_close(cpu(A.sht_matrix(60, az, co)), O.sht_matrix(60, az, co), rtol=1e-10, floor=4e-15)
    kr = np.array([1e-8, 1e-3, 250.0, 900.0])
    assert_close(cpu(R.weights(6, kr, "card")), O.weights(6, kr, "card"), rtol=1e-10, floor=1e-14)
