"""Angular (spherical / circular harmonic) matrices on the GPU.

Mirror of ``micarray/modal/angular.py`` of the reference: same names, argument
meaning, output shapes and exceptions; outputs are torch CUDA complex128.
The grid generators are O(Q) host-side setup and stay on the host (NumPy).
"""
import numpy as np
import torch

from .. import _lib
from ..util import asarray_1d, _device_of


def sht_matrix(N, azi, colat, weights=None):
    """Matrix of spherical harmonics up to order N (reference: angular.py:12-70).

    Returns ``Ymn`` of shape (Q, (N+1)**2), column ``n*n+n+m`` holding
    ``weights * Y_n^m(azi, colat)`` (Condon-Shortley phase, m = -n..n).
    A scalar ``colat`` broadcasts (angular.py:68); a scalar ``azi`` with a vector
    ``colat`` raises ValueError as in the reference.
    """
    lib = _lib.require_cuda()
    dev = _device_of(azi, colat, weights)
    azi = asarray_1d(azi, dev)
    colat = asarray_1d(colat, dev)
    Q = azi.numel()
    if colat.numel() not in (1, Q):
        raise ValueError("could not broadcast input array from shape ({},) into shape ({},)".format(
            colat.numel(), Q))
    w = None
    if weights is not None:
        w = torch.as_tensor(np.asarray(weights) if not isinstance(weights, torch.Tensor) else weights)
        w = w.to(device=dev, dtype=torch.float64).reshape(-1)
        if w.numel() == 1:
            w = w.expand(Q)
        elif w.numel() != Q:
            raise ValueError("operands could not be broadcast together with shapes ({},) ({},)".format(
                w.numel(), Q))
        w = w.contiguous()
    N = int(N)
    Y = torch.empty((Q, (N + 1) ** 2), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_sht_matrix(N, Q, _lib.ptr(azi), _lib.ptr(colat), int(colat.numel() == 1 and Q > 1),
                                      _lib.ptr(w), _lib.ptr(Y), _lib.stream_ptr(dev)))
    return Y


def cht_matrix(N, pol, weights=None):
    """Matrix of circular harmonics up to order N (reference: angular.py:106-150).

    Returns ``Psi`` of shape (Q, 2N+1) with column order ``[0, 1..N, -N..-1]``
    (angular.py:147).
    """
    lib = _lib.require_cuda()
    dev = _device_of(pol, weights)
    pol = asarray_1d(pol, dev)
    Q = pol.numel()
    w = None
    if weights is not None:
        w = torch.as_tensor(np.asarray(weights) if not isinstance(weights, torch.Tensor) else weights)
        w = w.to(device=dev, dtype=torch.float64).reshape(-1)
        if w.numel() == 1:
            w = w.expand(Q)
        elif w.numel() != Q:
            raise ValueError("operands could not be broadcast together with shapes ({},) ({},)".format(
                w.numel(), Q))
        w = w.contiguous()
    N = int(N)
    Psi = torch.empty((Q, 2 * N + 1), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_cht_matrix(N, Q, _lib.ptr(pol), _lib.ptr(w), _lib.ptr(Psi), _lib.stream_ptr(dev)))
    return Psi


def Legendre_matrix(N, ctheta):
    """Matrix of weighted Legendre polynomials ``(2n+1)/(4 pi) P_n(ctheta)``, shape (N+1, Q), complex
    (reference: angular.py:73-103, which evaluates the coefficient form with np.polyval; here the
    stable three-term recurrence)."""
    lib = _lib.require_cuda()
    dev = _device_of(ctheta)
    x = asarray_1d(ctheta, dev)
    Q = x.numel()
    N = int(N)
    L = torch.empty((N + 1, Q), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_legendre_matrix(N, Q, _lib.ptr(x), _lib.ptr(L), _lib.stream_ptr(dev)))
    return L


# ---------------------------------------------------------------------------
# Sampling grids: host-side, run once, O(Q) (reference: angular.py:153-235).
# ---------------------------------------------------------------------------
def grid_equal_angle(n):
    """Equi-angular sampling points on a sphere (reference: angular.py:153-185)."""
    m = 2 * n + 2
    azi = np.linspace(0, 2 * np.pi, m, endpoint=False)
    colat, step = np.linspace(0, np.pi, m, endpoint=False, retstep=True)
    colat = colat + step / 2
    p = np.arange(1, m, 2)
    w = 2 * np.pi / (n + 1) * np.sin(colat) * np.sum(np.sin(p[None, :] * colat[:, None]) / p[None, :], axis=1)
    return np.tile(azi, m), np.repeat(colat, m), np.repeat(w, m) / (n + 1)


def grid_gauss(n):
    """Gauss-Legendre sampling points on a sphere (reference: angular.py:188-214)."""
    azi = np.linspace(0, 2 * np.pi, 2 * n + 2, endpoint=False)
    x, w = np.polynomial.legendre.leggauss(n + 1)
    return np.tile(azi, n + 1), np.repeat(np.arccos(x), 2 * n + 2), np.repeat(w, 2 * n + 2) * np.pi / (n + 1)


def grid_equal_polar_angle(n):
    """Equi-angular sampling points on a circle (reference: angular.py:217-235)."""
    num_mic = 2 * n + 1
    return np.linspace(0, 2 * np.pi, num=num_mic, endpoint=False), np.ones(num_mic) / num_mic


def grid_lebedev(n):
    """Lebedev sampling points on a sphere (reference: angular.py:238-279).  Like the reference this
    needs the optional `quadpy` package (the reference imports it lazily, angular.py:6-9, and fails with
    a NameError without it; here the failure is an ImportError that says so).  Orders above 65 raise the
    reference's ValueError."""
    if n > 65:
        raise ValueError("Maximum available Lebedev grid order is 65. "
                         "(requested: {})".format(n))
    try:
        import quadpy
    except ImportError as e:
        raise ImportError("grid_lebedev needs the optional 'quadpy' package, as in the reference") from e
    from warnings import warn
    degrees = list(range(1, 32, 2)) + list(range(35, 132, 6))
    q = quadpy.sphere.Lebedev(degree=[x for x in degrees if x >= 2 * n][0])
    if np.any(q.weights < 0):
        warn("Lebedev grid of order {} has negative weights.".format(n))
    return q.azimuthal_polar[:, 0], q.azimuthal_polar[:, 1], 4 * np.pi * q.weights
