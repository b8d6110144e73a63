"""NumPy/SciPy restatement of the reference hot path (test infrastructure).

Every function names the reference lines it follows (paths relative to
``/root/reference``).  The special-function arithmetic of the reference lives
in a third-party dependency that is NOT vendored in the reference tree:
**SciPy ``scipy.special``** (declared un-pinned as ``'scipy'`` in
``setup.py:38-41``; this image has SciPy 1.18.1).  The oracle therefore calls
the same SciPy entry points as the reference's call sites do
(``angular.py:68``; ``radial.py:108,145-157,388,423-435``), with one
difference: ``scipy.special.sph_harm`` was removed from SciPy >= 1.17, so the
oracle calls its replacement ``sph_harm_y(n, m, colat, azi)``, which is what
the old function forwarded to.

The restatement is vectorised (one ufunc call over the whole grid) where the
reference loops in Python; the per-element arithmetic is the same ufunc, so the
results are bit-identical -- ``tests/test_oracle_golden.py`` checks that against
fixtures produced by the unmodified reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
--impl reference) may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import special

SETUPS = ("open", "card", "rigid")
METHODS = ("none", "discard", "hardclip", "softclip", "Tikh", "wng")


# --------------------------------------------------------------------------
# util
# --------------------------------------------------------------------------
def asarray_1d(a):
    """micarray/util.py:48-62 -- squeeze, scalar -> (1,), >1-D -> ValueError."""
    out = np.squeeze(np.asarray(a))
    if out.ndim == 0:
        return out.reshape((1,))
    if out.ndim > 1:
        raise ValueError("array must be one-dimensional")
    return out


# --------------------------------------------------------------------------
# stage 1: angular matrices
# --------------------------------------------------------------------------
def sh_index(N):
    """Column order of angular.py:65-69: i = n^2 + n + m, m = -n..n."""
    n = np.concatenate([np.full(2 * k + 1, k) for k in range(N + 1)])
    m = np.concatenate([np.arange(-k, k + 1) for k in range(N + 1)])
    return n, m


def sht_matrix(N, azi, colat, weights=None):
    """angular.py:12-70.  ``Ymn[q, n^2+n+m] = w_q * Y_n^m(azi_q, colat_q)``.

    A scalar ``colat`` broadcasts against a vector ``azi`` (angular.py:68);
    a scalar ``azi`` with vector ``colat`` raises ValueError because the
    reference assigns a (Qc,) vector into a (1,) column (angular.py:58-64,68).
    """
    azi = asarray_1d(azi)
    colat = asarray_1d(colat)
    Q = len(azi)
    if len(colat) not in (1, Q):
        raise ValueError(
            "could not broadcast input array from shape ({},) into shape ({},)".format(
                len(colat), Q))
    w = np.ones(Q) if weights is None else np.asarray(weights)
    n, m = sh_index(N)
    Y = special.sph_harm_y(n[None, :], m[None, :], colat[:, None], azi[:, None])
    Y = Y * np.reshape(w, (-1, 1)) if np.ndim(w) else Y * w
    return np.ascontiguousarray(np.broadcast_to(Y, (Q, (N + 1) ** 2)).astype(complex))


def ch_order(N):
    """angular.py:147 -- roll(arange(-N, N+1), -N) = [0, 1..N, -N..-1]."""
    return np.concatenate([np.arange(0, N + 1), np.arange(-N, 0)])


def cht_matrix(N, pol, weights=None):
    """angular.py:106-150.  ``Psi[q, i] = w_q * exp(1j * order[i] * pol_q)``."""
    pol = asarray_1d(pol)
    Q = len(pol)
    w = np.ones(Q) if weights is None else np.asarray(weights)
    order = ch_order(N)
    E = np.exp(1j * order[None, :] * pol[:, None])
    return (np.reshape(w, (-1, 1)) * E if np.ndim(w) else w * E).astype(complex)


# --------------------------------------------------------------------------
# stage 2: radial functions
# --------------------------------------------------------------------------
def _check_setup(setup):
    if setup not in SETUPS:
        raise ValueError("setup must be either: open, card or rigid")


def weights_2d(N, kr, setup):
    """radial.py:114-162 without the final squeeze: always (F, N+1).

    open  : j_n                                   (:146-147)
    card  : j_n - 1j * j_n'                       (:148-149)
    rigid : j_n - j_n'/h_n' * h_n, h = j - 1j*y   (:150-158); x == 0 -> j_n (:151-153)
    """
    _check_setup(setup)
    kr = asarray_1d(kr).astype(float)
    n = np.arange(N + 1)[None, :]
    x = kr[:, None]
    jn = special.spherical_jn(n, x)
    if setup == "open":
        return jn.astype(complex)
    jnd = special.spherical_jn(n, x, derivative=True)
    if setup == "card":
        return jn - 1j * jnd
    with np.errstate(all="ignore"):
        hn = jn - 1j * special.spherical_yn(n, x)
        hnd = jnd - 1j * special.spherical_yn(n, x, derivative=True)
        bn = jn - jnd / hnd * hn
    return np.where(x == 0, jn.astype(complex), bn)


def weights(N, kr, setup):
    """radial.py:114-162 including ``np.squeeze`` (:162)."""
    return np.squeeze(weights_2d(N, kr, setup))


def replace_zeros_of_radial_function(A, kr):
    """radial.py:165-213.

    Rows are sorted/deduplicated by ``np.unique(kr)`` (:196); every entry with
    ``abs < 1e-300`` (:198) takes the value of the entry in the same column at
    the nearest ``kr`` whose ORIGINAL value was non-zero (the ``zeros`` mask is
    not updated while replacing, :200-210; ``argmin`` breaks ties towards the
    lower sorted index); all-zero column -> ValueError (:209); rows are then
    un-permuted (:212) and the result squeezed (:213).
    """
    kr = asarray_1d(kr)
    A = np.array(A, copy=True)
    if A.ndim == 1 and A.shape[0] == len(kr):
        A = A[:, None]
    elif A.ndim == 1 and len(kr) == 1:
        A = A[None, :]
    if A.shape[0] != len(kr):
        raise ValueError("A and kr must have same len > 1,"
                         " but have {} and {}".format(A.shape[0], len(kr)))
    ukr, first, inverse = np.unique(kr, return_index=True, return_inverse=True)
    A = A[first, :]
    zero = np.abs(A) < 1e-300
    src = A.copy()
    ukr_f = ukr.astype(float)
    for j in np.flatnonzero(zero.any(axis=0)):
        good = ~zero[:, j]
        if not good.any():
            raise ValueError("Could not replace zero value in A.")
        for i in np.flatnonzero(zero[:, j]):
            dist = np.where(good, np.abs(ukr_f - ukr_f[i]), np.inf)
            A[i, j] = src[np.argmin(dist), j]
    return np.squeeze(A[inverse, :])


def spherical_pw(N, k, r, setup, replace_zeros=True):
    """radial.py:36-67 (+ decorator :9-33): ``4*pi * 1j**n * b_n(k*r)``."""
    kr = asarray_1d(k * r)
    n = np.arange(N + 1)
    out = 4 * np.pi * (1j) ** n * weights(N, kr, setup)
    if replace_zeros:
        return replace_zeros_of_radial_function(out, k)
    return out


def spherical_ps(N, k, r, rs, setup, replace_zeros=True):
    """radial.py:70-111: ``4*pi*(-1j) * k * h_n^(2)(k*rs) * b_n(k*r)``.

    Product order follows :109 (``bn * 4*pi * (-1j) * hn * k``) so that the
    k == 0 row becomes NaN exactly as in the reference.
    """
    k = asarray_1d(k)
    n = np.arange(N + 1)[None, :]
    bn = weights_2d(N, k * r, setup)
    x = (k * rs)[:, None]
    with np.errstate(all="ignore"):
        hn = special.spherical_jn(n, x) - 1j * special.spherical_yn(n, x)
        out = bn * 4 * np.pi * (-1j) * hn * k[:, None]
    out = np.squeeze(out)
    if replace_zeros:
        return replace_zeros_of_radial_function(out, k)
    return out


def circ_radial_weights_2d(N, kr, setup):
    """radial.py:393-441 without the squeeze: (F, 2N+1), orders [0..N, -N..-1].

    Orders 0..N from J_n, J_n', H_n^(2), H_n^(2)' (:423-436); negative orders
    are the (-1)^n mirror (:440).
    """
    _check_setup(setup)
    kr = asarray_1d(kr).astype(float)
    n = np.arange(N + 1)[None, :]
    x = kr[:, None]
    Jn = special.jv(n, x)
    if setup == "open":
        B = Jn.astype(complex)
    elif setup == "card":
        B = Jn - 1j * special.jvp(n, x, n=1)
    else:
        with np.errstate(all="ignore"):
            Jnd = special.jvp(n, x, n=1)
            Hn = special.hankel2(n, x)
            Hnd = special.h2vp(n, x)
            B = np.where(x == 0, Jn.astype(complex), Jn - Jnd / Hnd * Hn)
    sign = (-1.0) ** np.arange(N + 1)
    return np.concatenate((B, (B * sign)[:, :0:-1]), axis=-1)


def circ_radial_weights(N, kr, setup):
    return np.squeeze(circ_radial_weights_2d(N, kr, setup))


def circular_pw(N, k, r, setup, replace_zeros=True):
    """radial.py:317-348: ``1j**n * b_n(k*r)``, n = [0..N, -N..-1]."""
    kr = asarray_1d(k * r)
    n = ch_order(N)
    out = 1j ** n * circ_radial_weights(N, kr, setup)
    if replace_zeros:
        return replace_zeros_of_radial_function(out, k)
    return out


def circular_ls(N, k, r, rs, setup, replace_zeros=True):
    """radial.py:351-390: ``-1j/4 * H_n^(2)(k*rs) * b_n(k*r)``."""
    k = asarray_1d(k)
    n = ch_order(N)[None, :]
    bn = circ_radial_weights_2d(N, k * r, setup)
    with np.errstate(all="ignore"):
        out = -1j / 4 * np.squeeze(bn * special.hankel2(n, (k * rs)[:, None]))
    if replace_zeros:
        return replace_zeros_of_radial_function(out, k)
    return out


def regularize(dn, a0, method):
    """radial.py:216-266.  Returns ``(dn * hn, hn)``.

    ``hn`` keeps dn's (complex) dtype for none/discard/hardclip (``ones_like``)
    and is real for softclip/Tikh/wng.  Tikh first maps a0 -> sqrt(a0/2)
    (:255) so a0 < 2 yields NaN.
    """
    dn = np.asarray(dn)
    mag = np.abs(dn)
    over = mag > a0
    with np.errstate(all="ignore"):
        if method == "none":
            hn = np.ones_like(dn)
        elif method == "discard":
            hn = np.ones_like(dn)
            hn[over] = 0
        elif method == "hardclip":
            hn = np.ones_like(dn)
            hn[over] = a0 / mag[over]
        elif method == "softclip":
            hn = 2 / np.pi * np.arctan(np.pi / 2 * (a0 / mag))
        elif method == "Tikh":
            a = np.sqrt(a0 / 2)
            s = np.sqrt(1 - 1 / (a ** 2))
            alpha = (1 - s) / (1 + s)
            hn = 1 / (1 + alpha ** 2 * mag ** 2)
        elif method == "wng":
            hn = 1 / (mag ** 2)
        else:
            raise ValueError('method must be either: none, ' +
                             'discard, hardclip, softclip, Tikh or wng')
        return dn * hn, hn


def repeat_n_m(v):
    """radial.py:294-314: order-n entry repeated 2n+1 times along the last axis."""
    v = np.asarray(v)
    reps = 2 * np.arange(v.shape[-1]) + 1
    return np.squeeze(np.repeat(v, reps, axis=-1))


def _stack_diag(b):
    if b.ndim == 1:
        b = b[None, :]
    K, M = b.shape
    out = np.zeros((K, M, M), dtype=complex)
    idx = np.arange(M)
    out[:, idx, idx] = b
    return np.squeeze(out)


def diagonal_mode_mat(bk):
    """radial.py:269-291: repeat_n_m then one dense diagonal matrix per bin."""
    return _stack_diag(np.asarray(repeat_n_m(bk)))


def circ_diagonal_mode_mat(bk):
    """radial.py:444-465: as above but WITHOUT repeat_n_m."""
    return _stack_diag(np.asarray(bk))


# --------------------------------------------------------------------------
# stage 3: plane-wave decomposition / steering (the example idiom)
# --------------------------------------------------------------------------
def pwd_literal(Y_look, D, Y_mic, p):
    """doc/examples/modal_beamforming_rigid_spherical_array.py:38-39.

    ``A = matmul(matmul(Y_q, D), conj(Y_p.T)); q = matmul(A, p[..., None])``
    with ``D`` the dense (F, M, M) stack from diagonal_mode_mat; ``p`` is
    (F, Q) or (F, Q, T).  Only feasible for small sizes.
    """
    A = np.matmul(np.matmul(Y_look, D), np.conj(Y_mic.T))
    X = p[:, :, None] if p.ndim == 2 else p
    return np.matmul(A, X)


def pwd_factored(Y_look, d_cols, Y_mic, X):
    """Same contraction without materialising the diagonal stack.

    ``q_f = Y_L @ (d_f[:, None] * (Y_Q^H @ X_f))`` where ``d_cols`` is the
    (F, M) per-column diagonal (``repeat_n_m(dn)`` for spheres, ``dn`` itself
    for circles; radial.py:284 vs :459-465).  Algebraically identical to
    :func:`pwd_literal` (differs by rounding only, ~1e-15); needed because the
    dense stack is 77 GB at the large configs.
    """
    d_cols = np.asarray(d_cols)
    if d_cols.ndim == 1:
        d_cols = d_cols[None, :]
    X = X[:, :, None] if X.ndim == 2 else X
    Z = np.matmul(np.conj(Y_mic.T)[None], X)            # (F, M, T)
    Z *= d_cols[:, :, None]
    return np.matmul(Y_look[None], Z)                   # (F, L, T)


# --------------------------------------------------------------------------
# deterministic synthetic inputs shared by the tests and the bench
# --------------------------------------------------------------------------
def fibonacci_grid(P):
    i = np.arange(P)
    colat = np.arccos(1 - (2 * i + 1) / P)
    azi = np.mod(np.pi * (1 + np.sqrt(5)) * i, 2 * np.pi)
    return azi, colat, np.full(P, 4 * np.pi / P)


def grid_equal_angle(n):
    """angular.py:153-185 (host-side setup, restated for input generation)."""
    m = 2 * n + 2
    azi = np.linspace(0, 2 * np.pi, m, endpoint=False)
    colat, step = np.linspace(0, np.pi, m, endpoint=False, retstep=True)
    colat = colat + step / 2
    p = np.arange(1, m, 2)
    w = 2 * np.pi / (n + 1) * np.sin(colat) * np.sum(
        np.sin(p[None, :] * colat[:, None]) / p[None, :], axis=1)
    return (np.tile(azi, m), np.repeat(colat, m), np.repeat(w, m) / (n + 1))


def grid_gauss(n):
    """angular.py:188-214."""
    azi = np.linspace(0, 2 * np.pi, 2 * n + 2, endpoint=False)
    x, w = np.polynomial.legendre.leggauss(n + 1)
    return (np.tile(azi, n + 1), np.repeat(np.arccos(x), 2 * n + 2),
            np.repeat(w, 2 * n + 2) * np.pi / (n + 1))


def wavenumbers(F, fs=48000.0, c=343.0):
    return 2 * np.pi * np.linspace(0, fs / 2, F, endpoint=False) / c


def synth_spectra(F, Q, T, seed=0, dtype=np.complex64):
    rng = np.random.default_rng(seed)
    X = np.empty((F, Q, T), dtype=dtype)
    real_t = np.float32 if dtype == np.complex64 else np.float64
    # generated bin by bin so the 0.5-1 GB headline input never needs a
    # float64 temporary of the full size
    for f in range(F):
        g = rng.standard_normal((2, Q, T), dtype=np.float32)
        X[f].real = (g[0] * np.float32(np.sqrt(0.5))).astype(real_t)
        X[f].imag = (g[1] * np.float32(np.sqrt(0.5))).astype(real_t)
    return X


# --------------------------------------------------------------------------
# "next" rows
# --------------------------------------------------------------------------
def Legendre_matrix(N, ctheta):
    """angular.py:73-103: ``(2n+1)/(4 pi) * polyval(legendre(n), ctheta)``, shape (N+1, Q), complex."""
    ctheta = np.asarray(ctheta)
    M = 1 if ctheta.ndim == 0 else len(ctheta)
    out = np.zeros((N + 1, M), dtype=complex)
    for n in range(N + 1):
        out[n, :] = (2 * n + 1) / (4 * np.pi) * np.polyval(special.legendre(n), ctheta)
    return out


def matdiagmul(A, b):
    """util.py:65-93: ``C[k] = A * b[k]``."""
    b = np.asarray(b)
    if b.ndim == 1:
        b = b[None, :]
    return np.asarray(A)[None, :, :] * b[:, None, :]


def norm_of_columns(A, p=2):
    """util.py:5-21."""
    return np.linalg.norm(np.asarray(A), ord=p, axis=0)


def coherence_of_columns(A):
    """util.py:24-45: max off-diagonal magnitude of the Gram matrix of the column-normalised A."""
    A = np.asarray(A)
    An = A / norm_of_columns(A)[None, :]
    G = An.conj().T @ An
    np.fill_diagonal(G, 0)
    return np.max(np.abs(G))


# --------------------------------------------------------------------------- time <-> frequency ends
def irfft_shift(q, n=None):
    """Back end of every example script: ``np.fft.fftshift(np.fft.irfft(q, axis=0), axes=0)``
    (doc/examples/modal_beamforming_rigid_spherical_array.py:40, ..._open_circular_array.py:40,
    ..._open_spherical_array.py:40, ..._rigid_circular_array.py:37).  NumPy's pocketfft is the third-party
    arithmetic here (numpy!=1.11.0, setup.py:39); `n` as in np.fft.irfft (default 2 (F - 1))."""
    return np.fft.fftshift(np.fft.irfft(q, n=n, axis=0), axes=0)


def irfft_axis0(q, n=None):
    return np.fft.irfft(q, n=n, axis=0)


def stft(x, nfft, hop, window=None, bins=None):
    """Front end matching `sfa_stft` (the reference has none -- its examples start from spectra):
    X[f, q, t] = rfft(window * x[q, t*hop : t*hop + nfft])[f], no padding, no scaling; [F, Q, T] layout of the
    steering operator's input (the `p` of the examples with a frame axis)."""
    x = np.atleast_2d(np.asarray(x))
    Q, S = x.shape
    T = 1 + (S - nfft) // hop
    win = np.ones(nfft) if window is None else np.asarray(window)
    idx = np.arange(nfft)[None, :] + hop * np.arange(T)[:, None]
    frames = x[:, idx] * win                                  # [Q, T, nfft]
    X = np.fft.rfft(frames, axis=2)                           # [Q, T, nfft/2+1]
    X = np.transpose(X, (2, 0, 1))
    return X if bins is None else X[:bins]
