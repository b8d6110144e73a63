"""Stage 3 at the FULL sizes of BASELINE.json's configs 1-5: the result is checked on a few bins
against the oracle (which only finishes in seconds for a few bins at these sizes) and through a
size-independent property (linearity in X).  Config 3 -- the headline -- is run through exactly the
code paths bench.py times (`test_cfg3_full_size`)."""
import numpy as np
import pytest

from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


def bin_err(q, ref):
    q = q.reshape(q.shape[0], -1)
    ref = ref.reshape(ref.shape[0], -1)
    return float(np.max(np.max(np.abs(q - ref), axis=1) / np.max(np.abs(ref), axis=1)))


CONFIGS = {
    # name: (kind, N, Q grid, F, L, T, r, setup, a0, method)
    "cfg1": ("sph", 4, "equal_angle", 512, 360, 256, 0.042, "rigid", 100.0, "softclip"),
    "cfg2": ("circ", 16, 32, 1024, 360, 256, 0.1, "open", 3000.0, "softclip"),
    "cfg4": ("sph", 16, "gauss", 1024, 2562, 64, 0.1, "card", 100.0, "Tikh"),
    "cfg5": ("sph", 32, 1089, 4096, 16384, 32, 0.2, "rigid", 100.0, "Tikh"),
}


@pytest.mark.parametrize("name,precision", [(n, "c64_fast") for n in CONFIGS] + [("cfg4", "c64"), ("cfg1", "c64")])
def test_full_size_config(name, precision):
    import torch
    import sfa_numpy_b200 as micarray
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    kind, N, grid, F, L, T, r, setup, a0, method = CONFIGS[name]
    if kind == "sph":
        if grid == "equal_angle":
            az, co, w = O.grid_equal_angle(N)
        elif grid == "gauss":
            az, co, w = O.grid_gauss(N)
        else:
            az, co, w = O.fibonacci_grid(grid)
        if name == "cfg1":
            azl, col = np.linspace(0, 2 * np.pi, L, endpoint=False), np.full(L, np.pi / 2)
        else:
            azl, col, _ = O.fibonacci_grid(L)
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        k = O.wavenumbers(F) if name != "cfg4" else np.geomspace(0.01, 50, F) / r
        _, dn, _ = R.radial_filters(N, k, r, setup, a0, method)
        Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        with np.errstate(all="ignore"):
            dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, r, setup), a0, method)
        dcols = O.repeat_n_m(dn_ref)
    else:
        pol = np.linspace(0, 2 * np.pi, grid, endpoint=False)
        poll = np.linspace(0, 2 * np.pi, L, endpoint=False)
        w = np.full(grid, 1.0 / grid)
        Y_p, Y_q = A.cht_matrix(N, pol, w), A.cht_matrix(N, poll)
        k = O.wavenumbers(F)
        _, dn, _ = R.radial_filters(N, k, r, setup, a0, method, kind="circular_pw")
        Yp_ref, Yq_ref = O.cht_matrix(N, pol, w), O.cht_matrix(N, poll)
        with np.errstate(all="ignore"):
            dn_ref, _ = O.regularize(1 / O.circular_pw(N, k, r, setup), a0, method)
        dcols = dn_ref
    Q = Y_p.shape[0]
    assert np.max(np.abs(cpu(dn) - dn_ref) / (np.abs(dn_ref) + 1e-300)) < 1e-9
    g = torch.Generator(device="cuda").manual_seed(11)
    X1 = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda", generator=g)
    X2 = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda", generator=g)
    plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=precision)
    q1 = plan(X1)
    assert tuple(q1.shape) == (F, L, T) and bool(torch.isfinite(torch.view_as_real(q1)).all())
    # spot bins against the oracle (factored form; dense reference chain is infeasible here)
    sel = [1, F // 2, F - 1]
    ref = O.pwd_factored(Yq_ref, dcols[sel], Yp_ref, cpu(X1[sel]).astype(np.complex128))
    assert bin_err(cpu(q1[sel]), ref) < 1e-5
    # linearity over the whole output
    q2 = plan(X2)
    q12 = plan(X1 + 2 * X2)
    num = float((q12 - (q1 + 2 * q2)).abs().max())
    assert num / float(q12.abs().max()) < 5e-6


# bins at the hand-over points of the 74-bin frequency chunks of the slab pipeline (first / last bin of
# several chunks, 73 | 74 and 147 | 148 across a slab switch, the last, short chunk 962..1023) -- the fused
# single-launch kernel has no chunks, but it is held to the same bins
CFG3_SPOT_BINS = [0, 1, 73, 74, 147, 148, 517, 961, 962, 1023]


@pytest.mark.parametrize("precision,fused", [("c64_fast", True), ("c64", True), ("c64_fast", False), ("c64", False)])
def test_cfg3_full_size(precision, fused):
    """BASELINE configs[2] exactly as bench.py runs it: rigid sphere N=8, 64 mics (grid_equal_angle(3)),
    F=1024, L=2562 Fibonacci look directions, T=938 frames, Tikhonov a0=100, through PwdPlan into a
    row-padded beam tensor.  `fused=False` switches the fused persistent kernel off so that the
    chunked slab pipeline (pwd_steering_pipelined: 14 chunks, side-stream build, two alternating slabs)
    is held to the same bar."""
    import torch
    import sfa_numpy_b200 as micarray
    from sfa_numpy_b200 import _lib
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    lib = _lib.load()
    N, F, L, T, r = 8, 1024, 2562, 938, 0.042
    az, co, w = O.grid_equal_angle(3)
    azl, col, _ = O.fibonacci_grid(L)
    k = O.wavenumbers(F)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    _, dn, _ = R.radial_filters(N, k, r, "rigid", 100.0, "Tikh")
    Q = Y_p.shape[0]
    assert Q == 64
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda", generator=g)
    out = S.empty_beams(F, L, T, torch.complex64, X.device)
    out.fill_(float("nan"))
    _lib.check(lib.sfa_set_option(b"no_fused", 0.0 if fused else 1.0))
    try:
        plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=precision)
        q = plan(X, out=out)
        torch.cuda.synchronize()
    finally:
        lib.sfa_set_option(b"no_fused", 0.0)
    assert q.data_ptr() == out.data_ptr() and tuple(q.shape) == (F, L, T)
    # every output written, none left at the NaN fill
    assert bool(torch.isfinite(torch.view_as_real(q)).all())
    Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, r, "rigid"), 100.0, "Tikh")
    sel = CFG3_SPOT_BINS
    Xs = cpu(X[sel]).astype(np.complex128)
    ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref)[sel], Yp_ref, Xs)
    got = cpu(q[sel])
    err = [float(np.max(np.abs(got[i] - ref[i])) / np.max(np.abs(ref[i]))) for i in range(len(sel))]
    assert max(err) < 1e-5, dict(zip(sel, err))
    # power map of the same bins against NumPy
    pm = cpu(S.power_map(out)[sel])
    pm_ref = np.mean(np.abs(ref) ** 2, axis=2)
    assert np.max(np.abs(pm - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5
    # checksum-of-checksums over ALL bins: sum_l q[f, l, t] = (sum_l A_f[l, :]) X_f  -- a [F, Q] x [F, Q, T]
    # contraction the oracle finishes in a blink -- catches a wrong or missing tile anywhere in the 20 GB
    colsum = torch.zeros((F, T), dtype=torch.complex128, device=X.device)
    for f0 in range(0, F, 64):
        colsum[f0:f0 + 64] = out[f0:f0 + 64].to(torch.complex128).sum(dim=1)
    sA = np.einsum("m,fm,qm->fq", Yq_ref.sum(axis=0), O.repeat_n_m(dn_ref), np.conj(Yp_ref))
    ref_sum = np.einsum("fq,fqt->ft", sA, cpu(X).astype(np.complex128))
    binmax = torch.stack([out[f0:f0 + 64].abs().amax(dim=(1, 2)) for f0 in range(0, F, 64)]).reshape(-1)
    scale = np.sqrt(L) * cpu(binmax).astype(np.float64)[:, None]
    assert np.max(np.abs(cpu(colsum) - ref_sum) / scale) < 1e-5
