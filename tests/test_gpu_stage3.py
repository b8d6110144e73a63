"""GPU parity of stage 3 (plane-wave decomposition / steering) through the C ABI.
Tolerance (north_star): FP32-class beam outputs <= 1e-5, measured normwise per frequency bin
(max|q - q_ref| / max|q_ref| over the bin) because look directions in a null make an
elementwise relative error meaningless; the FP64 path is held to 1e-11."""
import ctypes

import numpy as np
import pytest

from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu
TOL64 = 1e-5
TOL128 = 1e-11


@pytest.fixture(scope="module")
def mic():
    import sfa_numpy_b200 as micarray
    return micarray


def cpu(t):
    return t.detach().cpu().numpy()


def bin_err(q, ref):
    q = q.reshape(q.shape[0], -1)
    ref = ref.reshape(ref.shape[0], -1)
    scale = np.max(np.abs(ref), axis=1)
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.max(np.abs(q - ref), axis=1) / scale))


def apply_err(q, ref, A, X):
    """Forward-error metric of the apply step q = A x: max|q - ref| / max(|A| |x|) per bin --
    the textbook scale of |fl(Ax) - Ax| <= gamma |A||x|.  Used where the beam output is the result
    of heavy cancellation (plane wave at kr ~ 0 with every order clipped to a0), where an error
    relative to max|q| measures the conditioning of the beamformer, not the arithmetic."""
    X = X[:, :, None] if X.ndim == 2 else X
    scale = np.max(np.matmul(np.abs(A), np.abs(X)).reshape(A.shape[0], -1), axis=1)
    err = np.max(np.abs(q - ref).reshape(A.shape[0], -1), axis=1)
    return float(np.max(err / np.where(scale > 0, scale, 1.0)))


def steering_matrix(Yq, dcols, Yp):
    return np.einsum("lm,fm,qm->flq", Yq, dcols, np.conj(Yp))


def test_cgemm3x_direct(mic):
    """out[b,n,m] = sum_k P[b,n,k] R[b,k,m] on tcgen05, ragged sizes, split-K, shared P, row scale."""
    import torch
    from sfa_numpy_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    worst = {0: 0.0, 1: 0.0}
    cases = [(1, 64, 128, 8), (2, 70, 200, 64), (3, 5, 33, 3), (1, 130, 129, 100), (2, 64, 256, 37),
             (1, 200, 1000, 64), (4, 17, 938, 9),
             # long inner dimension (in-kernel K loop when P is shared): folded bins, ragged n / k tails
             (3, 200, 40, 289), (2, 300, 64, 1089), (5, 129, 33, 65), (7, 16, 32, 97), (1, 257, 300, 130),
             # single (partial) row tile; exactly one accumulation block (9 x 32) and one k beyond it; many blocks
             (1, 40, 50, 100), (2, 64, 32, 288), (2, 64, 32, 289), (1, 130, 32, 2048)]
    for (B, Nn, Mm, K) in cases:
        for shared in (False, True):
            Pn = (rng.standard_normal((1 if shared else B, Nn, K)) + 1j * rng.standard_normal((1 if shared else B, Nn, K))).astype(np.complex64)
            Rn = (rng.standard_normal((B, K, Mm)) + 1j * rng.standard_normal((B, K, Mm))).astype(np.complex64)
            sc = (rng.standard_normal((B, Nn)) + 1j * rng.standard_normal((B, Nn))).astype(np.complex64)
            P = torch.from_numpy(Pn).cuda()
            R = torch.from_numpy(Rn).cuda()
            S = torch.from_numpy(sc).cuda()
            out = torch.full((B, Nn, Mm), float("nan"), dtype=torch.complex64, device="cuda")
            for use_scale, hybrid in ((False, 0), (True, 0), (False, 1), (True, 1)):
                _lib.check(lib.sfa_cgemm3x(B, Nn, Mm, K, _lib.ptr(P), K, 0 if shared else Nn * K, _lib.ptr(R), Mm,
                                           K * Mm, _lib.ptr(out), Mm, Nn * Mm, _lib.ptr(S) if use_scale else None,
                                           Nn, hybrid, _lib.stream_ptr()))
                ref = np.matmul(Pn.astype(np.complex128), Rn.astype(np.complex128))
                if use_scale:
                    ref = ref * sc[:, :, None].astype(np.complex128)
                got = cpu(out)
                assert np.isfinite(got.view(np.float32)).all(), (B, Nn, Mm, K, shared)
                err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
                worst[hybrid] = max(worst[hybrid], err)
                assert err < 5e-6, "cgemm3x %s shared=%s scale=%s hybrid=%d err %.3e" % (
                    (B, Nn, Mm, K), shared, use_scale, hybrid, err)
    print("cgemm3x worst normwise error: 3xTF32 %.2e, TF32+BF16 %.2e" % (worst[0], worst[1]))


@pytest.mark.parametrize("precision,tol", [("c128", TOL128), ("c64", TOL64), ("c64_fast", TOL64)])
@pytest.mark.parametrize("form", ["steering", "factored", "auto"])
def test_pwd_golden(mic, golden, precision, tol, form):
    g = golden["pwd"]
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    N = 4
    Y_p = A.sht_matrix(N, g["sph_azi"], g["sph_colat"], g["sph_w"])
    Y_q = A.sht_matrix(N, g["sph_azi_l"], np.pi / 2)
    bn = R.spherical_pw(N, g["sph_k"], 0.042, "rigid")
    dn, _ = R.regularize(1 / bn, 100, "softclip")
    q = S.pwd(Y_q, dn, Y_p, g["sph_X"], precision=precision, form=form)
    assert tuple(q.shape) == g["sph_q"].shape
    assert bin_err(cpu(q), g["sph_q"]) < tol
    # the literal reference chain on the GPU tensors (dense diagonal stack) gives the same
    if precision == "c128":
        import torch
        D = R.diagonal_mode_mat(dn)
        Alit = torch.matmul(torch.matmul(Y_q, D), Y_p.conj().T)
        assert bin_err(cpu(Alit), g["sph_A"]) < 1e-12
    N = 6
    Psi_p = A.cht_matrix(N, g["circ_pol"], g["circ_w"])
    Psi_q = A.cht_matrix(N, g["circ_pol_l"])
    Bn = R.circular_pw(N, g["sph_k"], 0.1, "open")
    Dn, _ = R.regularize(1 / Bn, 3000, "softclip")
    q = S.pwd(Psi_q, Dn, Psi_p, g["circ_X"], precision=precision, form=form)
    assert bin_err(cpu(q), g["circ_q"]) < tol


def _dot_sph(v, u):
    return (np.cos(v[0]) * np.sin(v[1]) * np.cos(u[0]) * np.sin(u[1]) +
            np.sin(v[0]) * np.sin(v[1]) * np.sin(u[0]) * np.sin(u[1]) + np.cos(v[1]) * np.cos(u[1]))


@pytest.mark.parametrize("precision,tol", [("c128", 1e-10), ("c64", TOL64), ("c64_fast", TOL64)])
def test_examples_end_to_end(mic, golden, precision, tol):
    """The four reference example scripts, re-run through the GPU API, against the q the
    unmodified scripts produced (fixtures) incl. the SURVEY 4.4 fingerprints."""
    g = golden["examples"]
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    k = g["open_spherical_k"]
    azi, colat, w = A.grid_gauss(20)
    azi_l = np.linspace(0, 2 * np.pi, 91, endpoint=False)
    Y_p = A.sht_matrix(20, azi, colat, w)
    Y_q = A.sht_matrix(20, azi_l, np.pi / 2)
    p = np.exp(1j * k[:, None] * 1 * _dot_sph((azi, colat), (np.pi, np.pi / 2)))
    for name, pp in (("open", p), ("rigid", g["rigid_spherical_p"])):
        dn, _ = R.regularize(1 / R.spherical_pw(20, k, 1, name), 100, "softclip")
        q = cpu(S.pwd(Y_q, dn, Y_p, pp, precision=precision))[:, :, 0]
        ref = g[name + "_spherical_q"]
        Aref = steering_matrix(cpu(Y_q), O.repeat_n_m(cpu(dn)), cpu(Y_p))
        assert apply_err(q, ref, Aref, pp) < tol, name
        # the plane wave arrives exactly between look directions 45 and 46 (a tie in exact arithmetic)
        assert int(np.argmax(np.abs(q[50]))) in (45, 46)
    pol, wc = A.grid_equal_polar_angle(30)
    pol_l = np.linspace(0, 2 * np.pi, 180, endpoint=False)
    Psi_p = A.cht_matrix(30, pol, wc)
    Psi_q = A.cht_matrix(30, pol_l)
    for name, a0, peak in (("open", 3000, 111), ("rigid", 100, 90)):
        Dn, _ = R.regularize(1 / R.circular_pw(30, k, 1, name), a0, "softclip")
        pp = g[name + "_circular_p"]
        q = cpu(S.pwd(Psi_q, Dn, Psi_p, pp, precision=precision))[:, :, 0]
        Aref = steering_matrix(cpu(Psi_q), cpu(Dn), cpu(Psi_p))
        assert apply_err(q, g[name + "_circular_q"], Aref, pp) < tol, name
        assert int(np.argmax(np.abs(q[50]))) == peak


SHAPES = [  # N, Q, L, F, T   (spherical unless N < 0 -> circular of order -N)
    (4, 50, 360, 16, 256),      # cfg 1 shape, fewer bins
    (4, 100, 360, 8, 1),        # Q > 64 -> split-K, T = 1
    (8, 64, 2562, 6, 130),      # cfg 3 shape, fewer bins / frames (ragged m tile)
    (8, 64, 65, 3, 938),        # ragged n tile, full T
    (-16, 32, 360, 32, 7),      # cfg 2 (circular), odd T
    (16, 289, 100, 4, 64),      # cfg 4: K = 289 (5 chunks)
    (3, 5, 7, 2, 3),            # tiny everything
    (2, 64, 128, 300, 128),     # many bins: several chunks
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precision,tol", [("c128", TOL128), ("c64", TOL64), ("c64_fast", TOL64)])
def test_pwd_shapes(mic, shape, precision, tol):
    N, Q, L, F, T = shape
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    rng = np.random.default_rng(abs(N) * 1000 + Q)
    k = O.wavenumbers(F) + 0.37
    if N >= 0:
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        dcols = O.repeat_n_m(dn_ref)
        dn, _ = R.regularize(1 / R.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
    else:
        N = -N
        pol = np.linspace(0, 2 * np.pi, Q, endpoint=False)
        poll = np.linspace(0, 2 * np.pi, L, endpoint=False)
        w = np.full(Q, 1 / Q)
        Yp_ref, Yq_ref = O.cht_matrix(N, pol, w), O.cht_matrix(N, poll)
        Y_p, Y_q = A.cht_matrix(N, pol, w), A.cht_matrix(N, poll)
        dn_ref, _ = O.regularize(1 / O.circular_pw(N, k, 0.1, "open"), 3000, "softclip")
        dcols = dn_ref
        dn, _ = R.regularize(1 / R.circular_pw(N, k, 0.1, "open"), 3000, "softclip")
    cd = np.complex128 if precision == "c128" else np.complex64
    X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(cd)
    ref = O.pwd_factored(Yq_ref, dcols.reshape(F, -1), Yp_ref, X.astype(np.complex128))
    for form in ("steering", "factored"):
        q = cpu(S.pwd(Y_q, dn, Y_p, X, precision=precision, form=form))
        assert q.shape == (F, L, T)
        assert bin_err(q, ref) < tol, (form, bin_err(q, ref))


def test_pwd_linearity_and_plan(mic):
    """Size-independent property at the headline shape (fewer bins): linear in X, and the
    reusable plan gives bit-identical results to the one-shot call."""
    import torch
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    N, Q, L, F, T = 8, 64, 2562, 40, 938
    az, co, w = O.grid_equal_angle(3)
    azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    k = O.wavenumbers(1024)[:F] + 1.0
    _, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
    g = torch.Generator(device="cuda").manual_seed(0)
    X1 = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda", generator=g)
    X2 = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda", generator=g)
    plan = S.PwdPlan(Y_q, dn, Y_p, T)
    q1, q2, q12 = plan(X1), plan(X2), plan(X1 + 2 * X2)
    scale = float(q12.abs().max())
    assert float((q12 - (q1 + 2 * q2)).abs().max()) / scale < 2e-6
    assert torch.equal(q1, S.pwd(Y_q, dn, Y_p, X1))
    # spot check 3 bins against the oracle
    Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
    sel = [0, 17, F - 1]
    ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref)[sel], Yp_ref, cpu(X1)[sel].astype(np.complex128))
    assert bin_err(cpu(q1)[sel], ref) < TOL64


def test_pwd_host_path(mic):
    """NumPy in / NumPy out through sfa_pwd_host_* (chunked H2D / compute / D2H pipeline)."""
    S = mic.modal.steering
    for T in (70, 138):                  # 138 > 128 and not a multiple of 32: row-padded device buffer + 2-D D2H copy
        _host_path_case(S, T)
    _host_path_case(S, 40, N=8, grid=100)    # 100 channels, 81 modes: factored form through the K-loop kernel
    # filters beyond the float32 range (an unregularised inverse near k = 0): the complex64 host path refuses them
    # (SFA_ERR_RANGE) instead of returning NaN beams; the complex128 path takes them
    from sfa_numpy_b200 import _lib
    az, co, w = O.grid_gauss(2)
    azl, col, _ = O.fibonacci_grid(50)
    Yp, Yq = O.sht_matrix(2, az, co, w), O.sht_matrix(2, azl, col)
    dn = np.ones((3, 3), dtype=complex)
    dn[1, 2] = 1e35
    with pytest.raises(_lib.SfaError, match="float32 range"):
        S.PwdHost(Yq, dn, Yp, 8, precision="c64")
    dn[1, 2] = np.inf
    with pytest.raises(_lib.SfaError, match="float32 range"):
        S.PwdHost(Yq, dn, Yp, 8, precision="c64_fast")
    dn[1, 2] = 1e35
    S.PwdHost(Yq, dn, Yp, 8, precision="c128").close()


def _host_path_case(S, T, N=4, grid=None):
    L, F = 200, 37
    az, co, w = O.grid_gauss(N) if grid is None else O.fibonacci_grid(grid)
    Q = len(az)
    azl, col, _ = O.fibonacci_grid(L)
    Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    k = O.wavenumbers(F) + 0.5
    dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "softclip")
    rng = np.random.default_rng(5)
    for precision, tol in (("c64", TOL64), ("c64_fast", TOL64), ("c128", TOL128)):
        cd = np.complex128 if precision == "c128" else np.complex64
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(cd)
        ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref), Yp_ref, X.astype(np.complex128))
        host = S.PwdHost(Yq_ref, dn_ref, Yp_ref, T, precision=precision, chunk_bins=8)
        Xp = host.pinned_empty(X.shape)
        Xp[...] = X
        q = host(Xp)
        assert bin_err(q, ref) < tol
        q2 = host(X)                     # pageable memory works too
        assert np.array_equal(q, q2)
        if precision != "c128":
            # power maps through the host path: with the beams, and alone (fused plans: map-only kernel runs)
            pm_ref = np.mean(np.abs(ref) ** 2, axis=2)
            q3 = np.empty_like(q)
            pm_a = host.power_map(Xp, out=q3)
            pm_b = host.power_map(Xp)
            assert np.array_equal(q3, q)
            for pm in (pm_a, pm_b):
                assert np.max(np.abs(pm - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5
        host.close()


@pytest.mark.parametrize("precision", ["c64", "c64_fast", "c128"])
def test_pwd_many_chunks(mic, precision):
    """Force a tiny per-chunk scratch budget so the bins are processed in many chunks (for the
    steering form this is the side-stream pipeline build(i+1) || gemm(i)); results must not change."""
    import torch
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    N, Q, L, F, T = 4, 36, 150, 90, 200
    az, co, w = O.fibonacci_grid(Q)
    azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    k = O.wavenumbers(F) + 0.3
    _, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "softclip")
    cd = torch.complex128 if precision == "c128" else torch.complex64
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((F, Q, T), dtype=cd, device="cuda", generator=g)
    from sfa_numpy_b200 import _lib
    lib = _lib.load()
    default_mb = lib.sfa_get_option(b"pwd_chunk_mb")
    try:
        for form in ("steering", "factored"):
            q_one = S.pwd(Y_q, dn, Y_p, X, precision=precision, form=form)
            _lib.check(lib.sfa_set_option(b"pwd_chunk_mb", 0.5))
            q_many = S.pwd(Y_q, dn, Y_p, X, precision=precision, form=form)
            _lib.check(lib.sfa_set_option(b"pwd_chunk_mb", default_mb))
            torch.cuda.synchronize()
            assert torch.equal(q_one, q_many), form
    finally:
        lib.sfa_set_option(b"pwd_chunk_mb", default_mb)
    Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "softclip")
    ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref), Yp_ref, cpu(X).astype(np.complex128))
    assert bin_err(cpu(q_many), ref) < (TOL128 if precision == "c128" else TOL64)


def test_pwd_degenerate(mic):
    import torch
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    az, co, w = O.fibonacci_grid(9)
    Y_p, Y_q = A.sht_matrix(2, az, co, w), A.sht_matrix(2, az[:4], co[:4])
    dn = torch.ones((0, 3), dtype=torch.complex128, device="cuda")
    q = S.pwd(Y_q, dn, Y_p, torch.zeros((0, 9, 5), dtype=torch.complex64, device="cuda"))
    assert tuple(q.shape) == (0, 4, 5)
    with pytest.raises(ValueError):
        S.pwd(Y_q, torch.ones((2, 4), dtype=torch.complex128, device="cuda"), Y_p,
              torch.zeros((2, 9, 5), dtype=torch.complex64, device="cuda"))          # 4 columns: neither M nor N+1
    with pytest.raises(ValueError):
        S.pwd(Y_q, torch.ones((2, 3), dtype=torch.complex128, device="cuda"), Y_p,
              torch.zeros((2, 8, 5), dtype=torch.complex64, device="cuda"))          # Q mismatch
    # 2-D p (T axis added as in np.matmul(A, p[:, :, None])) and a single bin
    dn1 = torch.tensor([[1.0, 2.0 + 1j, -0.5j]], dtype=torch.complex128, device="cuda")
    p = torch.randn((1, 9), dtype=torch.complex128, device="cuda")
    q = S.pwd(Y_q, dn1, Y_p, p)
    ref = O.pwd_factored(cpu(Y_q), O.repeat_n_m(cpu(dn1))[None, :], cpu(Y_p), cpu(p))
    assert tuple(q.shape) == (1, 4, 1) and bin_err(cpu(q), ref) < TOL128


@pytest.mark.parametrize("precision", ["c64", "c64_fast"])
def test_pwd_fused_kernel_shapes(mic, precision):
    """The fused persistent steering kernel (A_f built in tensor memory, pwd_fused_sm100.cu) against the
    oracle and against the slab build + GEMM pipeline it replaces, on ragged shapes: fewer than 128 look
    directions, one row more than a tile (ghost CTA of the pair), odd tile counts, frame counts around
    the 16/32/64-frame store boxes, few microphones (k padding), per-order and per-column filters."""
    import torch
    from sfa_numpy_b200 import _lib
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    lib = _lib.load()
    rng = np.random.default_rng(11)
    #        N  Q   L    F   T
    cases = [(2, 9, 5, 3, 2), (3, 16, 128, 2, 64), (3, 20, 129, 5, 70), (4, 36, 300, 7, 130), (2, 12, 700, 4, 18),
             (5, 64, 385, 3, 200), (7, 64, 2562, 2, 90), (1, 5, 257, 9, 34), (8, 64, 130, 3, 938),
             # L mod 128 = 16 (CUDA-core tail kernel, two row groups) and 17 (ragged tensor-core tile + ghost CTA)
             (3, 20, 144, 2, 40), (3, 20, 145, 2, 40), (3, 17, 266, 3, 333),
             # the highest orders of the Legendre build (16 groups: FP32 recurrence up to P_15), near-polar look directions
             (15, 64, 300, 3, 100), (12, 50, 1000, 2, 64)]
    for (N, Q, L, F, T) in cases:
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = O.wavenumbers(64)[:F] + 0.7
        dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref), Yp_ref, X.astype(np.complex128))
        Xd = torch.from_numpy(X).cuda()
        q_f = S.pwd(Yq_ref, dn_ref, Yp_ref, Xd, precision=precision, form="steering")
        assert bin_err(cpu(q_f), ref) < TOL64, ("fused", N, Q, L, F, T, bin_err(cpu(q_f), ref))
        # spherical-harmonic operators: the builders take the addition-theorem (Legendre) path; the same call with
        # the group-product table forced must agree with the oracle too -- and differ in the last bits, which
        # shows that the two calls did take different paths
        try:
            _lib.check(lib.sfa_set_option(b"no_legendre", 1.0))
            q_g = S.pwd(Yq_ref, dn_ref, Yp_ref, Xd, precision=precision, form="steering")
        finally:
            lib.sfa_set_option(b"no_legendre", 0.0)
        assert bin_err(cpu(q_g), ref) < TOL64, ("fused, table", N, Q, L, F, T, bin_err(cpu(q_g), ref))
        assert not torch.equal(q_f, q_g), ("Legendre path not taken", N, Q, L, F, T)
        try:
            _lib.check(lib.sfa_set_option(b"no_fused", 1.0))
            q_s = S.pwd(Yq_ref, dn_ref, Yp_ref, Xd, precision=precision, form="steering")
        finally:
            lib.sfa_set_option(b"no_fused", 0.0)
        assert bin_err(cpu(q_s), ref) < TOL64
        assert bin_err(cpu(q_f), cpu(q_s).astype(np.complex128)) < TOL64
        # plan API: prepared operators reused over calls, power map out of the kernel's epilogue (fused) or from
        # the extra pass (slab pipeline), both against NumPy
        pm_ref = np.mean(np.abs(ref) ** 2, axis=2)
        for nofused in (0.0, 1.0):
            try:
                _lib.check(lib.sfa_set_option(b"no_fused", nofused))
                plan = S.PwdPlan(Yq_ref, dn_ref, Yp_ref, T, precision=precision, form="steering")
                q1 = plan(Xd, power=True)
                q2 = plan(Xd, power=True)                    # second call: no re-preparation
            finally:
                lib.sfa_set_option(b"no_fused", 0.0)
            assert torch.equal(q1, q2) and bin_err(cpu(q1), ref) < TOL64
            pm = cpu(plan.power)
            assert pm.shape == pm_ref.shape
            assert np.max(np.abs(pm - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5, (N, Q, L, F, T, nofused)
    # per-column filters (circular array: d has M = 2N+1 columns, every column its own group): 13 groups (unrolled
    # builder), 33 groups = BASELINE config 2 (N = 16, 32 mics, 360 look directions; builder loop over the groups),
    # the latter also with enough bins for the stream-ahead conversion
    for (N, Q, L, F, T) in ((6, 13, 200, 6, 96), (16, 32, 360, 12, 256), (16, 32, 360, 200, 70), (31, 64, 150, 5, 40)):
        pol = np.linspace(0, 2 * np.pi, Q, endpoint=False)
        poll = np.linspace(0, 2 * np.pi, L, endpoint=False)
        Pp, Pq = O.cht_matrix(N, pol, np.full(Q, 1.0 / Q)), O.cht_matrix(N, poll)
        k = np.linspace(0.5, 40.0, F)
        Dn, _ = O.regularize(1 / O.circular_pw(N, k, 0.1, "open"), 3000, "softclip")
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        sel = sorted(set([0, F // 2, F - 1]))
        ref = O.pwd_factored(Pq, Dn[sel], Pp, X[sel].astype(np.complex128))
        plan = S.PwdPlan(Pq, Dn, Pp, T, precision=precision, form="steering")
        q = plan(torch.from_numpy(X).cuda(), power=True)
        assert bin_err(cpu(q[sel]), ref) < TOL64, (N, Q, L, F, T, bin_err(cpu(q[sel]), ref))
        # circular-harmonic operators: rotation build; the table build must agree with the oracle as well and differ
        # in the last bits (= the two runs took different paths)
        try:
            _lib.check(lib.sfa_set_option(b"no_legendre", 1.0))
            q_tab = S.PwdPlan(Pq, Dn, Pp, T, precision=precision, form="steering")(torch.from_numpy(X).cuda())
        finally:
            lib.sfa_set_option(b"no_legendre", 0.0)
        assert bin_err(cpu(q_tab[sel]), ref) < TOL64, ("table", N, Q, L, F, T, bin_err(cpu(q_tab[sel]), ref))
        assert not torch.equal(q, q_tab), ("rotation path not taken", N, Q, L, F, T)
        pm_ref = np.mean(np.abs(ref) ** 2, axis=2)
        assert np.max(np.abs(cpu(plan.power)[sel] - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5
        q_auto = S.pwd(Pq, Dn, Pp, torch.from_numpy(X).cuda(), precision=precision)       # form chosen by the plan
        assert bin_err(cpu(q_auto[sel]), ref) < TOL64


@pytest.mark.parametrize("precision", ["c64_fast", "c64"])
def test_pwd_fused_stream_ahead_conversion(mic, precision):
    """Launches with many bins: the fused kernel's spare warps split the spectra of the bins ahead of the ones
    being multiplied (hand-over through L2 with release/acquire counters, tail look directions and their power
    sums on the way).  Same arithmetic as the pre-pass, so the beams must be BIT-identical to a run with
    `fused_prepass` = 1 (every bin split by the pre-pass kernel), and both must match the oracle on spot bins.
    Shapes: ghost CTA in the pair (odd tile count), tail rows (L mod 128 <= 16), fewer 32-frame tiles than
    slices (idle slices), few microphones (padded k), the headline's row geometry."""
    import torch
    from sfa_numpy_b200 import _lib
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    lib = _lib.load()
    rng = np.random.default_rng(5)
    #        N  Q   L    F    T
    cases = [(3, 20, 389, 260, 100), (2, 9, 700, 200, 45), (4, 25, 2562, 64, 130), (3, 16, 256, 300, 33),
             (2, 12, 130, 600, 64)]
    for (N, Q, L, F, T) in cases:
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp_ref, Yq_ref = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = np.linspace(0.5, 40.0, F)
        dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        Xd = torch.from_numpy(X).cuda()
        res = {}
        for prepass in (0.0, 1.0):
            try:
                _lib.check(lib.sfa_set_option(b"fused_prepass", prepass))
                plan = S.PwdPlan(Yq_ref, dn_ref, Yp_ref, T, precision=precision, form="steering")
                q1 = plan(Xd, power=True).clone()
                pm1 = plan.power.clone()
                for _ in range(3):                           # repeated launches reuse counters and planes
                    q2 = plan(Xd, power=True)
                assert torch.equal(q1, q2) and torch.equal(pm1, plan.power), (N, Q, L, F, T, prepass)
                res[prepass] = (q1, pm1)
            finally:
                lib.sfa_set_option(b"fused_prepass", 0.0)
        assert torch.equal(res[0.0][0], res[1.0][0]), ("beams differ from the pre-pass run", N, Q, L, F, T)
        pm_a, pm_b = cpu(res[0.0][1]), cpu(res[1.0][1])
        assert np.max(np.abs(pm_a - pm_b) / np.max(pm_b, axis=1, keepdims=True)) < 1e-6
        bins = sorted(set([0, 1, F // 3, F // 2, F - 2, F - 1]))
        ref = O.pwd_factored(Yq_ref, O.repeat_n_m(dn_ref)[bins], Yp_ref, X[bins].astype(np.complex128))
        got = cpu(res[0.0][0][bins])
        assert bin_err(got, ref) < TOL64, (N, Q, L, F, T, bin_err(got, ref))
        pm_ref = np.mean(np.abs(ref) ** 2, axis=2)
        assert np.max(np.abs(pm_a[bins] - pm_ref) / np.max(pm_ref, axis=1, keepdims=True)) < 1e-5


def test_pwd_power_map_only(mic):
    """``plan(X, power=True, beams=False)``: the fused kernel skips its stores -- the map must equal the one of a run
    that also writes the beams, bit for bit; plans that are not the fused kernel refuse."""
    import torch
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    rng = np.random.default_rng(9)
    for (N, Q, L, F, T) in ((3, 20, 389, 260, 100), (4, 36, 300, 7, 130)):
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp, Yq = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = np.linspace(0.5, 40.0, F)
        dn, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = torch.from_numpy((rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)).cuda()
        plan = S.PwdPlan(Yq, dn, Yp, T, precision="c64_fast", form="steering")
        q = plan(X, power=True)
        pm_full = plan.power.clone()
        plan.power.zero_()
        pm = plan(X, power=True, beams=False)
        assert pm is plan.power and torch.equal(pm, pm_full)
        ref = np.mean(np.abs(cpu(q).astype(np.complex128)) ** 2, axis=2)
        assert np.max(np.abs(cpu(pm) - ref) / np.max(ref, axis=1, keepdims=True)) < 1e-5
        with pytest.raises(ValueError):
            plan(X, beams=False)
    plan = S.PwdPlan(Yq, dn, Yp, T, precision="c64_fast", form="factored")
    with pytest.raises(ValueError):
        plan(X, power=True, beams=False)


@pytest.mark.parametrize("precision", ["c64_fast", "c64"])
def test_pwd_row_padding_contract(mic, precision):
    """sfa_pwd_shape.out_pad_writable: with the flag (the library's own row-padded tensors, `empty_beams`) the fused
    kernel completes the last 128-byte line of every row -- frames [T, roundup(T, 16)) become zeros, nothing
    beyond is touched; without it (a caller's view of a wider tensor) nothing beyond frame T changes.  The beams
    themselves are bit-equal either way."""
    import torch
    from sfa_numpy_b200 import _lib
    lib = _lib.load()
    A, R, S = mic.modal.angular, mic.modal.radial, mic.modal.steering
    rng = np.random.default_rng(5)
    for (N, Q, L, F, T) in ((3, 16, 300, 3, 138), (4, 25, 130, 40, 200), (8, 64, 2562, 2, 938), (2, 9, 200, 2, 131),
                            (2, 9, 130, 3, 193), (3, 16, 128, 2, 257)):
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp, Yq = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = O.wavenumbers(128)[:F] + 0.5
        dn, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        Xd = torch.from_numpy(X).cuda()
        ld, T16 = S._padded_pitch(T), (T + 15) // 16 * 16
        assert ld >= T16 > T
        plan = S.PwdPlan(Yq, dn, Yp, T, precision=precision, form="steering")
        # (1) a caller's view: the sentinel beyond frame T survives (also through the slab build + GEMM pipeline)
        big = torch.full((F, L, ld), 7.0 + 3.0j, dtype=torch.complex64, device="cuda")
        q_view = plan(Xd, out=big[:, :, :T])
        assert bool((big[:, :, T:] == 7.0 + 3.0j).all()), (N, Q, L, F, T)
        for form in ("steering", "factored"):
            try:
                _lib.check(lib.sfa_set_option(b"no_fused", 1.0))
                big2 = torch.full((F, L, ld), 7.0 + 3.0j, dtype=torch.complex64, device="cuda")
                q2 = S.pwd(Yq, dn, Yp, Xd, precision=precision, form=form, out=big2[:, :, :T])
            finally:
                lib.sfa_set_option(b"no_fused", 0.0)
            assert bool((big2[:, :, T:] == 7.0 + 3.0j).all()), ("no_fused", form, N, Q, L, F, T)
            assert bin_err(cpu(q2), cpu(q_view).astype(np.complex128)) < TOL64
        # (2) the library's own padded tensor: zeros up to the next multiple of 16 frames, the rest untouched
        own = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
        base = own._base if own._base is not None else own
        base.fill_(7.0 + 3.0j)
        q_own = plan(Xd, out=own)
        main = (L // 128) * 128 if (L % 128 and L % 128 <= 16 and L > 128) else L        # tail rows: plain stores
        assert bool((base[:, :main, T:T16] == 0).all()), (N, Q, L, F, T)
        assert bool((base[:, :, T16:] == 7.0 + 3.0j).all()) and bool((base[:, main:, T:] == 7.0 + 3.0j).all())
        assert torch.equal(q_own, q_view)
        ref = O.pwd_factored(Yq, O.repeat_n_m(dn), Yp, X.astype(np.complex128))
        assert bin_err(cpu(q_own), ref) < TOL64
        # (3) a slice of an owned tensor does not carry the mark unless the caller says so
        sl = own[:, :, :]
        assert not S._pad_owned(sl) and S._pad_owned(own)


@pytest.mark.parametrize("precision", ["c64_fast", "c64"])
def test_pwd_fused_operator_structure_detection(mic, precision):
    """The addition-theorem build of the fused kernel is taken only when the preparation has verified, in FP64, that
    every group product follows s_q (2n+1) P_n(x_lq).  Operators that are not spherical-harmonic matrices --
    random complex matrices with per-order filters, SH matrices with one perturbed entry, SH matrices with rows of
    Y_look rescaled -- must run through the group-product table: correct against the oracle and BIT-equal to a run
    with the Legendre path switched off.  SH matrices with complex weights folded into Y_mic still qualify."""
    import torch
    from sfa_numpy_b200 import _lib
    S = mic.modal.steering
    lib = _lib.load()
    rng = np.random.default_rng(23)

    def both(Yq, dn, Yp, Xd):
        q1 = S.pwd(Yq, dn, Yp, Xd, precision=precision, form="steering")
        try:
            _lib.check(lib.sfa_set_option(b"no_legendre", 1.0))
            q2 = S.pwd(Yq, dn, Yp, Xd, precision=precision, form="steering")
        finally:
            lib.sfa_set_option(b"no_legendre", 0.0)
        return q1, q2

    for (N, Q, L, F, T) in ((3, 16, 300, 4, 96), (8, 64, 515, 3, 130), (1, 4, 140, 2, 70)):
        M = (N + 1) ** 2
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp_sh, Yq_sh = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        dn = rng.standard_normal((F, N + 1)) + 1j * rng.standard_normal((F, N + 1))
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        Xd = torch.from_numpy(X).cuda()
        Yq_rand = rng.standard_normal((L, M)) + 1j * rng.standard_normal((L, M))
        Yp_rand = rng.standard_normal((Q, M)) + 1j * rng.standard_normal((Q, M))
        Yq_pert = Yq_sh.copy()
        Yq_pert[L // 2, M - 1] *= 1.0 + 1e-6
        Yq_rows = Yq_sh * (1.0 + 0.5 * rng.random((L, 1)))
        for name, Yq, Yp in (("random", Yq_rand, Yp_rand), ("perturbed", Yq_pert, Yp_sh), ("row-scaled", Yq_rows, Yp_sh)):
            ref = O.pwd_factored(Yq, O.repeat_n_m(dn), Yp, X.astype(np.complex128))
            q1, q2 = both(Yq, dn, Yp, Xd)
            assert bin_err(cpu(q1), ref) < TOL64, (name, N, Q, L, F, T, bin_err(cpu(q1), ref))
            assert torch.equal(q1, q2), (name, "took the Legendre path", N, Q, L, F, T)
        # per-column filters (one group per column): random operators, and circular-harmonic matrices whose filters
        # are NOT symmetric in +-m (the rotation build must not assume d_m = d_-m) or with one perturbed entry
        Nc = min(2 * N + 3, 15)
        Mc = 2 * Nc + 1
        Qc = min(Q, 24)
        pol = np.linspace(0, 2 * np.pi, Qc, endpoint=False) + 0.1
        poll = np.sort(rng.random(L)) * 2 * np.pi
        Pp, Pq = O.cht_matrix(Nc, pol, np.full(Qc, 1.0 / Qc)), O.cht_matrix(Nc, poll)
        dcol = rng.standard_normal((F, Mc)) + 1j * rng.standard_normal((F, Mc))
        Xc = (rng.standard_normal((F, Qc, T)) + 1j * rng.standard_normal((F, Qc, T))).astype(np.complex64)
        Xcd = torch.from_numpy(Xc).cuda()
        ref = O.pwd_factored(Pq, dcol, Pp, Xc.astype(np.complex128))
        q1, q2 = both(Pq, dcol, Pp, Xcd)
        assert bin_err(cpu(q1), ref) < TOL64 and bin_err(cpu(q2), ref) < TOL64, ("circular", bin_err(cpu(q1), ref))
        assert not torch.equal(q1, q2), ("circular: rotation path not taken", N, Q, L, F, T)
        Pq_pert = Pq.copy()
        Pq_pert[L // 3, Mc - 2] *= 1.0 + 1e-6
        Pq_rand = rng.standard_normal((L, Mc)) + 1j * rng.standard_normal((L, Mc))
        for name, Yq_c in (("circular perturbed", Pq_pert), ("circular random", Pq_rand)):
            ref = O.pwd_factored(Yq_c, dcol, Pp, Xc.astype(np.complex128))
            q1, q2 = both(Yq_c, dcol, Pp, Xcd)
            assert bin_err(cpu(q1), ref) < TOL64, (name, bin_err(cpu(q1), ref))
            assert torch.equal(q1, q2), (name, "took the rotation path")
        # complex column weights on the microphone side keep the structure (s_q becomes complex)
        Yp_cw = Yp_sh * np.exp(1j * rng.random((Q, 1)))
        ref = O.pwd_factored(Yq_sh, O.repeat_n_m(dn), Yp_cw, X.astype(np.complex128))
        q1, q2 = both(Yq_sh, dn, Yp_cw, Xd)
        assert bin_err(cpu(q1), ref) < TOL64 and bin_err(cpu(q2), ref) < TOL64
        assert not torch.equal(q1, q2), ("complex weights: Legendre path not taken", N, Q, L, F, T)


@pytest.mark.parametrize("precision", ["c64_fast", "c64"])
def test_pwd_fused_pair_mode(mic, precision):
    """The two ways the CTAs of a cluster share one bin: as ONE tcgen05.mma.cta_group::2 pair (M = 256, every CTA
    holding half of the frame tile) and as two MMA streams behind a multicast plane stream.  Same MMA sequence per
    output element, so the results must be BIT-equal -- on ragged shapes: odd tile counts (ghost CTA of a pair),
    frame tails around the 16/32/64-frame boxes, few microphones, the CUDA-core tail rows, power map, map-only
    runs, enough bins for the stream-ahead conversion."""
    import torch
    from sfa_numpy_b200 import _lib
    S = mic.modal.steering
    lib = _lib.load()
    lib.sfa_get_option.restype = ctypes.c_double
    default_pair = lib.sfa_get_option(b"fused_pair")
    rng = np.random.default_rng(31)
    #        N  Q   L    F   T
    cases = [(3, 16, 256, 2, 64), (3, 20, 300, 5, 70), (4, 36, 385, 7, 130), (2, 12, 700, 4, 18), (7, 64, 2562, 2, 90),
             (1, 5, 257, 9, 34), (8, 64, 640, 3, 938), (3, 17, 266, 3, 333), (4, 25, 515, 70, 161)]
    for (N, Q, L, F, T) in cases:
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp, Yq = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = O.wavenumbers(128)[:F] + 0.7
        dn, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = (rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)
        Xd = torch.from_numpy(X).cuda()
        res = {}
        for pair in (0.0, 1.0):
            try:
                _lib.check(lib.sfa_set_option(b"fused_pair", pair))
                plan = S.PwdPlan(Yq, dn, Yp, T, precision=precision, form="steering")
                q = plan(Xd, power=True).clone()
                assert lib.sfa_pwd_is_fused(ctypes.byref(plan.shape)), ("not the fused kernel", N, Q, L, F, T)
                pm = plan.power.clone()
                pm_only = plan(Xd, power=torch.empty_like(pm), beams=False).clone()
                q2 = plan(Xd, power=True)
                assert torch.equal(q, q2)
                res[pair] = (q, pm, pm_only)
            finally:
                lib.sfa_set_option(b"fused_pair", default_pair)
        sel = sorted(set([0, F // 2, F - 1]))
        ref = O.pwd_factored(Yq, O.repeat_n_m(dn)[sel], Yp, X[sel].astype(np.complex128))
        for pair in (0.0, 1.0):
            assert bin_err(cpu(res[pair][0][sel]), ref) < TOL64, (pair, N, Q, L, F, T, bin_err(cpu(res[pair][0][sel]), ref))
            assert torch.equal(res[pair][1], res[pair][2]), ("map-only", pair, N, Q, L, F, T)
        assert torch.equal(res[0.0][0], res[1.0][0]), ("pair vs multicast", N, Q, L, F, T)
        assert torch.equal(res[0.0][1], res[1.0][1])
