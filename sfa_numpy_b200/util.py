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
