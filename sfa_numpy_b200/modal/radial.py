"""Radial filters on the GPU.

Mirror of ``micarray/modal/radial.py`` of the reference: same names, argument
meaning, squeeze rules (survey A1) and ValueErrors; outputs are torch CUDA
tensors (complex128 / float64).
"""
import numpy as np
import torch

from .. import _lib
from ..util import asarray_1d, as_complex_tensor, _device_of

_SETUPS = {"open": 0, "card": 1, "rigid": 2}
_METHODS = {"none": 0, "discard": 1, "hardclip": 2, "softclip": 3, "Tikh": 4, "wng": 5}
_KINDS = {"weights": 0, "spherical_pw": 1, "spherical_ps": 2,
          "circ_radial_weights": 3, "circular_pw": 4, "circular_ls": 5}


def _setup_id(setup):
    if setup not in _SETUPS:
        raise ValueError('setup must be either: open, card or rigid')       # radial.py:160,438
    return _SETUPS[setup]


def _method_id(method):
    if method not in _METHODS:
        raise ValueError('method must be either: none, ' +
                         'discard, hardclip, softclip, Tikh or wng')        # radial.py:263-264
    return _METHODS[method]


class _DeferredChecks:
    """Zero-replacement status words copied to pinned host memory without stalling the stream.
    `flush()` (called by the next radial call and by `radial.check_deferred()`) raises the
    ValueError the reference would have raised (radial.py:209)."""

    def __init__(self):
        self.pending = []
        self.pool = []

    def add(self, status):
        # pinned buffers are recycled: cudaHostAlloc per call would cost more than the kernels
        host = self.pool.pop() if self.pool else torch.empty(4, dtype=torch.int32, pin_memory=True)
        # copy and event go to the stream of the tensor's OWN device (which need not be the current one)
        with torch.cuda.device(status.device):
            host.copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(status.device))
        self.pending.append((ev, host))

    def flush(self, wait=False):
        """Validate every status word that has landed (all of them with wait=True).  An entry leaves the
        pending list BEFORE it can raise, so one failed call is reported exactly once; entries that are
        not ready yet are re-queued; if several failed, the first is raised and the rest stay queued."""
        pending, self.pending = self.pending, []
        error = None
        for ev, host in pending:
            if error is not None or (not wait and not ev.query()):
                self.pending.append((ev, host))
                continue
            ev.synchronize()
            st = host.tolist()
            self.pool.append(host)
            try:
                _raise_for_status(st)
            except (ValueError, RuntimeError) as e:
                error = e
        if error is not None:
            raise error


_deferred = _DeferredChecks()


def check_deferred():
    """Wait for and validate every deferred zero-replacement status (see `check="defer"`)."""
    _deferred.flush(wait=True)


def _raise_for_status(st):
    if st[1]:
        raise ValueError("Could not replace zero value in A.")           # radial.py:209
    if st[2]:
        raise RuntimeError("sfa_radial_filter: more than 2**20 zero entries to replace")


def _radial(kind, N, k, r, rs, setup, replace_zeros, method=None, a0=0.0, check="now"):
    """One launch of sfa_radial_filter.  Returns (bn, dn, hn) un-squeezed ([F, C])."""
    lib = _lib.require_cuda()
    sid = _setup_id(setup)
    mid = -1 if method is None else _method_id(method)
    dev = k.device
    N = int(N)
    F = k.numel()
    C = (N + 1) if kind < 3 else (2 * N + 1)
    bn = torch.empty((F, C), dtype=torch.complex128, device=dev)
    dn = hn = None
    if mid >= 0:
        dn = torch.empty((F, C), dtype=torch.complex128, device=dev)
        hn = torch.empty((F, C), dtype=torch.complex128 if mid <= 2 else torch.float64, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    work = None
    if replace_zeros:
        work = torch.empty(lib.sfa_radial_work_bytes(kind, N, F), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_radial_filter(kind, N, F, _lib.ptr(k), float(r), float(rs), sid,
                                         int(bool(replace_zeros)), mid, float(a0), _lib.ptr(bn), _lib.ptr(dn),
                                         _lib.ptr(hn), _lib.ptr(status), _lib.ptr(work), _lib.stream_ptr(dev)))
    if replace_zeros:
        if check == "none":
            pass                                      # graph capture / caller reads `status` itself
        elif check == "defer":
            # queue this call's status FIRST: an error of an earlier call raised by flush() must not
            # lose it (the kernel has already been launched and its result is valid)
            _deferred.add(status)
            _deferred.flush()
        else:
            _raise_for_status(status.tolist())
    return bn, dn, hn


def _as_wavenumber(k, r):
    """k*r exactly as the reference forms it (radial.py:63,344): a Python list times a float raises
    TypeError there, so it does here."""
    if isinstance(k, torch.Tensor):
        return asarray_1d(k)
    if isinstance(k, (list, tuple)):
        k * r if not isinstance(r, int) else None        # noqa: B018  (TypeError as in the reference)
    return asarray_1d(k)


def weights(N, kr, setup):
    """Radial weighting functions b_n(kr) (reference: radial.py:114-162)."""
    kr = asarray_1d(kr)
    bn, _, _ = _radial(_KINDS["weights"], N, kr, 1.0, 0.0, setup, False)
    return bn.squeeze()


def spherical_pw(N, k, r, setup, replace_zeros=True):
    """Radial coefficients for a plane wave, ``4 pi i^n b_n(kr)`` (reference: radial.py:36-67)."""
    k = _as_wavenumber(k, r)
    bn, _, _ = _radial(_KINDS["spherical_pw"], N, k, r, 0.0, setup, replace_zeros)
    return bn.squeeze()


def spherical_ps(N, k, r, rs, setup, replace_zeros=True):
    """Radial coefficients for a point source (reference: radial.py:70-111)."""
    k = asarray_1d(k)
    bn, _, _ = _radial(_KINDS["spherical_ps"], N, k, r, rs, setup, replace_zeros)
    return bn.squeeze()


def circ_radial_weights(N, kr, setup):
    """Circular radial weighting functions (reference: radial.py:393-441)."""
    kr = asarray_1d(kr)
    bn, _, _ = _radial(_KINDS["circ_radial_weights"], N, kr, 1.0, 0.0, setup, False)
    return bn.squeeze()


def circular_pw(N, k, r, setup, replace_zeros=True):
    """Radial coefficients for a plane wave on a circle, ``i^n b_n(kr)`` (reference: radial.py:317-348)."""
    k = _as_wavenumber(k, r)
    bn, _, _ = _radial(_KINDS["circular_pw"], N, k, r, 0.0, setup, replace_zeros)
    return bn.squeeze()


def circular_ls(N, k, r, rs, setup, replace_zeros=True):
    """Radial coefficients for a line source (reference: radial.py:351-390)."""
    k = asarray_1d(k)
    bn, _, _ = _radial(_KINDS["circular_ls"], N, k, r, rs, setup, replace_zeros)
    return bn.squeeze()


def radial_filters(N, k, r, setup, a0, method, kind="spherical_pw", rs=0.0, replace_zeros=True, check="now"):
    """Fused ``bn = <kind>(N, k, r, setup); dn, hn = regularize(1/bn, a0, method)`` in ONE kernel
    pass (the inverse never makes a separate trip through HBM).  Returns (bn, dn, hn), squeezed.
    Equivalent reference sequence: modal_beamforming_rigid_spherical_array.py:34-35.
    ``check="defer"`` keeps the call asynchronous: the "Could not replace zero value" check is
    validated later (next radial call or ``check_deferred()``) instead of synchronising now;
    ``check="none"`` skips it (CUDA-graph capture: the same inputs were validated once before)."""
    if kind not in _KINDS:
        raise ValueError("unknown radial function: %r" % (kind,))
    kk = asarray_1d(k)
    bn, dn, hn = _radial(_KINDS[kind], N, kk, r, rs, setup,
                         replace_zeros and kind not in ("weights", "circ_radial_weights"), method, a0, check)
    return bn.squeeze(), dn.squeeze(), hn.squeeze()


def replace_zeros_of_radial_function(A, kr):
    """Replace zero entries A[i, j] with A[l, j] != 0 at the nearest kr[l] (reference: radial.py:165-213)."""
    lib = _lib.require_cuda()
    dev = _device_of(A, kr)
    kr = asarray_1d(kr, dev)
    A = as_complex_tensor(A, dev)
    if A.ndim == 1 and A.shape[0] == kr.numel():
        A = A[:, None]
    elif A.ndim == 1 and kr.numel() == 1:
        A = A[None, :]
    if A.shape[0] != kr.numel():
        raise ValueError("A and kr must have same len > 1,"
                         " but have {} and {}".format(A.shape[0], kr.numel()))
    # np.unique(kr, True, True) (radial.py:196): sorted distinct kr, first occurrence, inverse map
    ukr, inverse = torch.unique(kr, sorted=True, return_inverse=True)
    first = torch.full((ukr.numel(),), kr.numel(), dtype=torch.long, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(kr.numel(), device=dev), reduce="amin")
    As = A[first, :].contiguous()
    F, C = As.shape
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    nbytes = lib.sfa_radial_work_bytes(0, 0, F)
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_replace_zeros(F, C, _lib.ptr(ukr), _lib.ptr(As), _lib.ptr(status), _lib.ptr(work),
                                         nbytes, _lib.stream_ptr(dev)))
    st = status.tolist()
    if st[1]:
        raise ValueError("Could not replace zero value in A.")
    if st[2]:
        raise RuntimeError("sfa_replace_zeros: more than 2**20 zero entries to replace")
    return As[inverse, :].squeeze()


def regularize(dn, a0, method):
    """Regularization (amplitude limitation) of radial filters (reference: radial.py:216-266).
    Returns ``(dn * hn, hn)``; hn is complex for none/discard/hardclip and real otherwise."""
    lib = _lib.require_cuda()
    mid = _method_id(method)
    dn = as_complex_tensor(dn)
    dev = dn.device
    out = torch.empty_like(dn)
    hn = torch.empty(dn.shape, dtype=torch.complex128 if mid <= 2 else torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_regularize(dn.numel(), _lib.ptr(dn), float(a0), mid, _lib.ptr(out), _lib.ptr(hn),
                                      _lib.stream_ptr(dev)))
    return out, hn


def repeat_n_m(v):
    """Repeat the order-n entry 2n+1 times along the last axis (reference: radial.py:294-314)."""
    lib = _lib.require_cuda()
    v = as_complex_tensor(v)
    dev = v.device
    if v.ndim > 2:
        raise ValueError("repeat_n_m expects a vector or a stack of vectors")
    v2 = v.reshape(1, -1) if v.ndim == 1 else v
    F, C = v2.shape
    out = torch.empty((F, C * C), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_repeat_n_m(F, C - 1, _lib.ptr(v2.contiguous()), _lib.ptr(out), _lib.stream_ptr(dev)))
    return out.squeeze()


def _diag_stack(b):
    lib = _lib.load()
    dev = b.device
    if b.ndim == 1:
        b = b[None, :]
    K, M = b.shape
    out = torch.empty((K, M, M), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_diagonal_stack(K, M, _lib.ptr(b.contiguous()), _lib.ptr(out), _lib.stream_ptr(dev)))
    return out.squeeze()


def diagonal_mode_mat(bk):
    """Dense (K, M, M) stack of diagonal matrices, M = (N+1)**2 (reference: radial.py:269-291).
    Materialises K*M*M complex128 values: meant for small sizes / drop-in compatibility; the
    steering operator (modal.steering.pwd) takes the vector ``bk`` directly."""
    _lib.require_cuda()
    return _diag_stack(repeat_n_m(bk))


def circ_diagonal_mode_mat(bk):
    """As diagonal_mode_mat but without repeat_n_m (reference: radial.py:444-465)."""
    _lib.require_cuda()
    return _diag_stack(as_complex_tensor(bk))
