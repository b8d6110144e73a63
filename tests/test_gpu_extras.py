"""GPU parity of the SURVEY 8f "next" rows: Legendre_matrix, util helpers, forward-model synthesis."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, floor=1e-15)


def cpu(t):
    return t.detach().cpu().numpy()


def test_legendre_and_util(golden):
    import sfa_numpy_b200 as micarray
    g = golden["extras"]
    L = micarray.modal.angular.Legendre_matrix(12, g["leg_ct"])
    assert tuple(L.shape) == (13, 41) and str(L.dtype) == "torch.complex128"
    # the reference's coefficient-form polyval loses digits at n = 12; 1e-9 is its own accuracy there
    assert_close(cpu(L), g["leg12"], rtol=1e-9, floor=1e-12, name="legendre")
    # against the exact recurrence (scipy eval_legendre)
    from scipy import special
    n = np.arange(41)[:, None]
    x = np.cos(np.linspace(0.01, 3.1, 300))
    ref = (2 * n + 1) / (4 * np.pi) * special.eval_legendre(n, x[None, :])
    assert_close(cpu(micarray.modal.angular.Legendre_matrix(40, x)).real, ref, rtol=1e-10, floor=1e-14, name="leg40")
    U = micarray.util
    assert_close(cpu(U.matdiagmul(g["mdm_A"], g["mdm_b"])), g["mdm_C"], **TOL)
    assert_close(cpu(U.matdiagmul(g["mdm_A"], g["mdm_b"][0])), g["mdm_C1"], **TOL)
    assert_close(cpu(U.norm_of_columns(g["mdm_A"])), g["noc2"], **TOL)
    assert_close(cpu(U.norm_of_columns(g["mdm_A"], p=1)), g["noc1"], **TOL)
    assert abs(float(U.coherence_of_columns(g["mdm_A"])) - float(g["coh"])) < 1e-12
    assert abs(float(U.coherence_of_columns(g["coh_Y_in"])) - float(g["coh_Y"])) < 1e-12
    rng = np.random.default_rng(1)
    A = rng.standard_normal((300, 70)) + 1j * rng.standard_normal((300, 70))
    assert abs(float(U.coherence_of_columns(A)) - O.coherence_of_columns(A)) < 1e-12
    assert micarray.util.db(np.ones(3), power=True) == 10         # reference's literal (survey A13)


def test_forward_model_synthesis(golden):
    """p = Y_p D conj(Y_pw.T) of the rigid-sphere example (rigid_spherical_array.py:21-26) on the GPU
    equals the p the unmodified script produced."""
    import sfa_numpy_b200 as micarray
    g = golden["examples"]
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    k = g["rigid_spherical_k"]
    azi, colat, _ = A.grid_gauss(20)
    bn = R.spherical_pw(20, k, 1, setup="rigid")
    Y_p = A.sht_matrix(20, azi, colat)
    Y_pw = A.sht_matrix(20, np.pi, np.pi / 2)
    p = S.synthesize(Y_p, bn, Y_pw)
    assert tuple(p.shape) == (100, 882, 1)
    ref = g["rigid_spherical_p"]
    assert np.max(np.abs(cpu(p)[:, :, 0] - ref)) < 1e-11 * np.max(np.abs(ref))
    # circular (open_circular_array.py:22-27, Nsf = 50)
    pol, _ = A.grid_equal_polar_angle(30)
    Bn = R.circular_pw(50, g["open_circular_k"], 1, setup="open")
    p = S.synthesize(A.cht_matrix(50, pol), Bn, A.cht_matrix(50, 1.23 * np.pi))
    ref = g["open_circular_p"]
    assert np.max(np.abs(cpu(p)[:, :, 0] - ref)) < 1e-11 * np.max(np.abs(ref))


def test_power_map():
    """mean_t |q|^2 of contiguous and row-padded beam tensors, odd and even T, against NumPy."""
    import torch
    import sfa_numpy_b200 as micarray
    S = micarray.modal.steering
    rng = np.random.default_rng(5)
    for (F, L, T) in ((3, 17, 938), (2, 5, 7), (1, 64, 128), (4, 9, 1)):
        qn = (rng.standard_normal((F, L, T)) + 1j * rng.standard_normal((F, L, T))).astype(np.complex64)
        ref = np.mean(np.abs(qn.astype(np.complex128)) ** 2, axis=-1)
        for padded in (False, True):
            q = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda")) if padded else \
                torch.empty((F, L, T), dtype=torch.complex64, device="cuda")
            q.copy_(torch.from_numpy(qn))
            pm = S.power_map(q).cpu().numpy()
            assert pm.shape == (F, L) and pm.dtype == np.float32
            assert np.max(np.abs(pm - ref) / ref) < 2e-6
    with pytest.raises(ValueError):
        S.power_map(torch.zeros((2, 2, 2), dtype=torch.complex64))


@pytest.mark.parametrize("n", [0, 1, 3, 8, 16, 32, 60])
def test_grids_on_device(n):
    """angular.grid_* with device= (sfa_grid kernel) against the reference's NumPy formulas
    (angular.py:153-235; oracle restatement, golden-pinned) -- FP64, 1e-13 absolute on angles and weights."""
    import torch
    import sfa_numpy_b200 as micarray
    A = micarray.modal.angular
    for name, ref in (("grid_equal_angle", O.grid_equal_angle(n)), ("grid_gauss", O.grid_gauss(n))):
        got = getattr(A, name)(n, device="cuda")
        assert len(got) == 3
        for g, r in zip(got, ref):
            assert g.is_cuda and g.dtype == torch.float64 and tuple(g.shape) == r.shape
            np.testing.assert_allclose(g.cpu().numpy(), r, rtol=0, atol=2e-13)
        assert abs(float(got[2].sum()) - 4 * np.pi) < 1e-11
    pol, w = A.grid_equal_polar_angle(n, device="cuda")
    num = 2 * n + 1
    np.testing.assert_array_equal(pol.cpu().numpy(), np.linspace(0, 2 * np.pi, num=num, endpoint=False))
    np.testing.assert_array_equal(w.cpu().numpy(), np.ones(num) / num)
    # the host flavour is unchanged and the device grids feed sht_matrix directly
    azi, colat, wt = A.grid_gauss(n, device="cuda")
    Y = A.sht_matrix(n, azi, colat, wt)
    Yh = O.sht_matrix(n, *O.grid_gauss(n))
    assert float(np.max(np.abs(Y.cpu().numpy() - Yh))) < 1e-12
    with pytest.raises(ValueError):
        A.grid_gauss(-1, device="cuda")
