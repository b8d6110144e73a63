"""Host-side input coercion mirroring micarray/util.py of the reference."""
import numpy as np
import torch


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("sfa_numpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return default_device()


def asarray_1d(a, device=None, dtype=torch.float64):
    """micarray/util.py:48-62: squeeze, scalar -> shape (1,), more than 1-D -> ValueError.
    Returns a contiguous 1-D tensor on `device`."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.as_tensor(np.asarray(a))
    t = t.squeeze()
    if t.ndim == 0:
        t = t.reshape(1)
    elif t.ndim > 1:
        raise ValueError("array must be one-dimensional")
    if device is None:
        device = _device_of(a)
    return t.to(device=device, dtype=dtype).contiguous()


def as_complex_tensor(a, device=None, dtype=torch.complex128):
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.as_tensor(np.asarray(a))
    if device is None:
        device = _device_of(a)
    return t.to(device=device, dtype=dtype).contiguous()


def db(x, power=False):
    """micarray/util.py:96-109 with the reference's (buggy) power=True behaviour kept:
    it returns the literal 10 (survey A13)."""
    if power:
        return 10
    x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return 20 * torch.log10(torch.abs(x))


def matdiagmul(A, b):
    """Multiplication of a matrix and (a stack of) diagonal matrices: ``C[k] = A * b[k]``
    (reference: micarray/util.py:65-93).  A: (M, N); b: (N,) or (K, N); returns (K, M, N)."""
    from . import _lib
    lib = _lib.require_cuda()
    dev = _device_of(A, b)
    A = as_complex_tensor(A, dev)
    b = as_complex_tensor(b, dev)
    if b.ndim == 1:
        b = b[None, :]
    K, N = b.shape
    M, N2 = A.shape
    if N2 != N:
        raise ValueError("operands could not be broadcast together with shapes ({},{}) ({},)".format(M, N2, N))
    C = torch.empty((K, M, N), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_matdiagmul(K, M, N, _lib.ptr(A), _lib.ptr(b), _lib.ptr(C), _lib.stream_ptr(dev)))
    return C


def norm_of_columns(A, p=2):
    """Vector p-norm of each column of a matrix (reference: micarray/util.py:5-21)."""
    from . import _lib
    lib = _lib.require_cuda()
    A = as_complex_tensor(A)
    dev = A.device
    M, N = A.shape
    out = torch.empty(N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_norm_of_columns(M, N, _lib.ptr(A), float(p), _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


def coherence_of_columns(A):
    """Mutual coherence of the columns of A (reference: micarray/util.py:24-45)."""
    from . import _lib
    lib = _lib.require_cuda()
    A = as_complex_tensor(A)
    dev = A.device
    M, N = A.shape
    norms = torch.empty(max(N, 1), dtype=torch.float64, device=dev)
    res = torch.zeros(1, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_coherence_of_columns(M, N, _lib.ptr(A), _lib.ptr(norms), _lib.ptr(res),
                                                _lib.stream_ptr(dev)))
    return res[0]
