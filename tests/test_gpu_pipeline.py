"""`pipeline.ModalBeamformer`: lines 18-39 of the reference's example scripts as one CUDA-graph launch
(doc/examples/modal_beamforming_rigid_spherical_array.py, ..._open_circular_array.py) against the oracle, graph
replay against eager launches, a bin slab (the multi-GPU use), the fused power map."""
import numpy as np
import pytest

from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu


def _err(q, ref):
    q = q.reshape(q.shape[0], -1)
    ref = ref.reshape(ref.shape[0], -1)
    return float(np.max(np.max(np.abs(q - ref), axis=1) / np.max(np.abs(ref), axis=1)))


def test_modal_beamformer_spherical():
    import torch
    from sfa_numpy_b200.pipeline import ModalBeamformer
    N, F, L, T = 4, 48, 300, 70
    az, co, w = O.grid_gauss(N)
    azl, col, _ = O.fibonacci_grid(L)
    k = O.wavenumbers(F) + 0.3
    X = O.synth_spectra(F, len(az), T, seed=4)
    Yp, Yq = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    dn, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100.0, "softclip")
    ref = O.pwd_factored(Yq, O.repeat_n_m(dn), Yp, X.astype(np.complex128))
    for bins in (None, (8, 40)):
        f0, f1 = (0, F) if bins is None else bins
        bf = ModalBeamformer(N, (az, co, w), (azl, col), k, 0.042, "rigid", 100.0, "softclip", T, precision="c64_fast",
                             bins=bins, power=True)
        assert bf.graph is not None, bf.graph_error
        q = bf(torch.from_numpy(X[f0:f1]).cuda())
        torch.cuda.synchronize()
        got = q.cpu().numpy().copy()
        assert got.shape == (f1 - f0, L, T) and _err(got, ref[f0:f1]) < 1e-5
        pm_ref = np.mean(np.abs(ref[f0:f1]) ** 2, axis=2)
        pm = bf.power.cpu().numpy()
        assert np.max(np.abs(pm - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5
        # replay == eager, and a second batch through the same graph
        q_e = bf.run_eager().clone()
        assert torch.equal(q_e, bf.run())
        X2 = O.synth_spectra(F, len(az), T, seed=5)
        ref2 = O.pwd_factored(Yq, O.repeat_n_m(dn)[f0:f1], Yp, X2[f0:f1].astype(np.complex128))
        assert _err(bf(torch.from_numpy(X2[f0:f1]).cuda()).cpu().numpy(), ref2) < 1e-5
        assert bf.launches_per_run >= 5
    with pytest.raises(ValueError):
        ModalBeamformer(N, (az, co, w), (azl, col), k, 0.042, "rigid", 100.0, "softclip", T, bins=(10, F + 1))
    with pytest.raises(ValueError):
        ModalBeamformer(N, (az, co, w), (azl, col), k, 0.042, "rigid", 100.0, "softclip", T, kind="toroidal")


def test_modal_beamformer_circular():
    import torch
    from sfa_numpy_b200.pipeline import ModalBeamformer
    N, Q, F, L, T = 6, 13, 20, 90, 33
    pol = np.linspace(0, 2 * np.pi, Q, endpoint=False)
    wq = np.full(Q, 1.0 / Q)
    poll = np.linspace(0, 2 * np.pi, L, endpoint=False)
    k = O.wavenumbers(F) + 0.4
    X = O.synth_spectra(F, Q, T, seed=6)
    Pp, Pq = O.cht_matrix(N, pol, wq), O.cht_matrix(N, poll)
    Dn, _ = O.regularize(1 / O.circular_pw(N, k, 0.1, "open"), 3000.0, "softclip")
    ref = O.pwd_factored(Pq, Dn, Pp, X.astype(np.complex128))
    for precision, tol in (("c64", 1e-5), ("c128", 1e-10)):
        bf = ModalBeamformer(N, (pol, wq), (poll,), k, 0.1, "open", 3000.0, "softclip", T, kind="circular", precision=precision)
        Xd = torch.from_numpy(X.astype(np.complex128 if precision == "c128" else np.complex64)).cuda()
        q = bf(Xd)
        torch.cuda.synchronize()
        assert _err(q.cpu().numpy(), ref) < tol, (precision, _err(q.cpu().numpy(), ref))
