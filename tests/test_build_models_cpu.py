"""CPU: the two structure models behind the fused steering kernel's builds (pwd_fused_sm100.cu), restated in NumPy
and checked against the oracle's matrices -- the identity the kernels rely on, the FP64 acceptance test of the
preparation, and the FP32 recurrences the builder warps run.

  spherical arrays (addition theorem):  sum_m Y_n^m(l) conj(Y_n^m(q)) w_q = s_q (2n+1) P_n(x_lq)
  circular arrays:                      Psi_L[l, c] conj(Psi_Q[q, c])      = s_q z_lq^m(c),  orders [0..N, -N..-1]
"""
import numpy as np
import pytest

from oracle import sfa_oracle as O

TOL_MODEL = 1e-8          # kLegTol of the kernel: largest deviation relative to the largest entry of the group


def group_products(Y_L, Y_Q, groups):
    return [np.einsum("lm,qm->lq", Y_L[:, g], np.conj(Y_Q[:, g])) for g in groups]


def sph_model(Y_L, Y_Q, N):
    """What group_products_g2_kernel computes in FP64: s_q, x_lq and the per-order deviation record."""
    groups = [np.arange(n * n, (n + 1) ** 2) for n in range(N + 1)]
    G = group_products(Y_L, Y_Q, groups)
    s = Y_L[0, 0] * np.conj(Y_Q[:, 0])
    den = 3.0 * np.abs(s) ** 2
    x = np.where(den > 0, np.real(G[1] * np.conj(s)[None, :]) / np.where(den > 0, den, 1.0), 0.0)
    x = np.clip(x, -1.0, 1.0)
    P = [np.ones_like(x), x]
    for n in range(1, N):
        P.append(((2 * n + 1) * x * P[n] - n * P[n - 1]) / (n + 1))
    dev = [np.max(np.abs(G[n] - s[None, :] * (2 * n + 1) * P[n])) for n in range(N + 1)]
    mag = [np.max(np.abs(G[n])) for n in range(N + 1)]
    return G, s, x, np.array(dev), np.array(mag)


@pytest.mark.parametrize("N", [1, 4, 8, 15])
def test_addition_theorem_model_and_fp32_recurrence(N):
    rng = np.random.default_rng(N)
    Q, L, F = min((N + 1) ** 2, 64), 300, 5
    az, co, w = O.fibonacci_grid(Q)
    azl, col, _ = O.fibonacci_grid(L)
    Y_Q, Y_L = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    G, s, x, dev, mag = sph_model(Y_L, Y_Q, N)
    assert np.all(dev <= 1e-12 * mag), (dev / mag)                   # far inside the kernel's 1e-8 acceptance
    # the builders' arithmetic: FP32 three-term recurrence on x, filters scaled by (2n+1), s_q applied at the end
    d = (rng.standard_normal((F, N + 1)) + 1j * rng.standard_normal((F, N + 1))) * 10.0 ** rng.uniform(-1, 2, (F, N + 1))
    A_ref = np.einsum("fn,nlq->flq", d, np.stack(G))
    x32 = x.astype(np.float32)
    dd = (d * (2 * np.arange(N + 1) + 1)).astype(np.complex64)
    p0, p1 = np.ones_like(x32), x32.copy()
    acc = dd[:, 0, None, None] + dd[:, 1, None, None] * x32[None]
    for n in range(1, N):
        ca, cb = np.float32((2 * n + 1) / (n + 1)), np.float32(n / (n + 1))
        pn = (ca * x32) * p1 - cb * p0
        p0, p1 = p1, pn.astype(np.float32)
        acc = (acc + dd[:, n + 1, None, None] * p1[None]).astype(np.complex64)
    A32 = acc * s.astype(np.complex64)[None, None, :]
    err = np.max(np.abs(A32 - A_ref), axis=(1, 2)) / np.max(np.abs(A_ref), axis=(1, 2))
    assert np.max(err) < 5e-6, err


def test_addition_theorem_model_rejects_other_operators():
    N, Q, L = 4, 25, 200
    az, co, w = O.fibonacci_grid(Q)
    azl, col, _ = O.fibonacci_grid(L)
    Y_Q, Y_L = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
    rng = np.random.default_rng(0)
    pert = Y_L.copy()
    pert[L // 2, -1] *= 1.0 + 1e-6
    rows = Y_L * (1.0 + 0.5 * rng.random((L, 1)))
    rand = rng.standard_normal(Y_L.shape) + 1j * rng.standard_normal(Y_L.shape)
    for name, YL in (("perturbed", pert), ("row-scaled", rows), ("random", rand)):
        _, _, _, dev, mag = sph_model(YL, Y_Q, N)
        assert np.any(dev > TOL_MODEL * mag), name
    # complex weights on the microphone side keep the structure
    _, _, _, dev, mag = sph_model(Y_L, Y_Q * np.exp(1j * rng.random((Q, 1))), N)
    assert np.all(dev <= 1e-12 * mag)


@pytest.mark.parametrize("N", [3, 16, 31])
def test_rotation_model_and_fp32_recurrence(N):
    rng = np.random.default_rng(100 + N)
    M, Q, L, F = 2 * N + 1, 32, 250, 4
    pol = np.linspace(0, 2 * np.pi, Q, endpoint=False) + 0.05
    poll = np.sort(rng.random(L)) * 2 * np.pi
    P_Q, P_L = O.cht_matrix(N, pol, np.full(Q, 1.0 / Q)), O.cht_matrix(N, poll)
    order = O.ch_order(N)
    assert list(order[:N + 1]) == list(range(N + 1)) and list(order[N + 1:]) == list(range(-N, 0))
    G = group_products(P_L, P_Q, [np.array([c]) for c in range(M)])
    s = P_L[0, 0] * np.conj(P_Q[:, 0])
    z = G[1] * np.conj(s)[None, :] / (np.abs(s) ** 2)[None, :]
    for c in range(M):
        m = c if c <= N else c - M
        model = s[None, :] * (z ** m if m >= 0 else np.conj(z) ** (-m))
        assert np.max(np.abs(G[c] - model)) <= 1e-12 * np.max(np.abs(G[c])), c
    # builders: filters NOT symmetric in +-m; (cos mD, sin mD) by repeated FP32 rotation
    d = rng.standard_normal((F, M)) + 1j * rng.standard_normal((F, M))
    A_ref = np.einsum("fc,clq->flq", d, np.stack(G))
    c32, s32 = np.real(z).astype(np.float32), np.imag(z).astype(np.float32)
    C, S = c32.copy(), s32.copy()
    d32 = d.astype(np.complex64)
    acc = np.broadcast_to(d32[:, 0, None, None], (F, L, Q)).astype(np.complex64)
    for m in range(1, N + 1):
        dp, dm = d32[:, m], d32[:, M - m]
        e, o = dp + dm, 1j * (dp - dm)
        acc = (acc + e[:, None, None] * C[None] + o[:, None, None] * S[None]).astype(np.complex64)
        C, S = (C * c32 - S * s32).astype(np.float32), (S * c32 + C * s32).astype(np.float32)
    A32 = acc * s.astype(np.complex64)[None, None, :]
    err = np.max(np.abs(A32 - A_ref), axis=(1, 2)) / np.max(np.abs(A_ref), axis=(1, 2))
    assert np.max(err) < 5e-6, err
