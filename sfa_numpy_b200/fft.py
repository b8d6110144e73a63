"""Time <-> frequency ends of the beamforming path on the GPU.

Every example script of the reference ends with

    q_t = np.fft.fftshift(np.fft.irfft(q, axis=0), axes=0)

(doc/examples/modal_beamforming_rigid_spherical_array.py:40, ..._open_circular_array.py:40,
..._open_spherical_array.py:40, ..._rigid_circular_array.py:37): `irfft_shift` is that line on a CUDA
tensor, `irfft` the plain ``np.fft.irfft(q, n, axis=0)``.  `stft` is the matching front end that
produces spectra ``X[F, Q, T]`` directly in the layout `modal.steering.pwd` reads.  Hand-written FFT
kernels behind the C ABI (`sfa_irfft_axis0`, `sfa_stft`); no cuFFT, no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .util import _device_of


def _as_cuda_complex(q):
    dev = _device_of(q)
    if isinstance(q, torch.Tensor):
        t = q.detach()
    else:
        t = torch.as_tensor(np.asarray(q))
    if not t.is_complex():
        t = t.to(torch.complex128 if t.dtype == torch.float64 else torch.complex64)
    return t.to(dev).contiguous()


def irfft(q, n=None, shift=False):
    """``np.fft.irfft(q, n, axis=0)`` (then ``np.fft.fftshift(., axes=0)`` when `shift`) for a complex
    array [F, ...]; complex64 -> float32, complex128 -> float64.  Returns a CUDA tensor [n, ...]."""
    lib = _lib.require_cuda()
    t = _as_cuda_complex(q)
    if t.ndim == 0:
        raise ValueError("irfft needs at least one axis")
    F = t.shape[0]
    n_out = 2 * (F - 1) if n is None else int(n)
    if n_out < 1:
        raise ValueError("Invalid number of FFT data points ({}) specified.".format(n_out))   # numpy's text
    tail = tuple(t.shape[1:])
    C = int(np.prod(tail)) if tail else 1
    is_double = t.dtype == torch.complex128
    out = torch.empty((n_out,) + tail, dtype=torch.float64 if is_double else torch.float32, device=t.device)
    if C == 0:
        return out
    nbytes = lib.sfa_irfft_work_bytes(n_out, int(is_double))
    work = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.sfa_irfft_axis0(F, C, n_out, _lib.ptr(t), C, _lib.ptr(out), C, int(is_double), int(bool(shift)),
                                       _lib.ptr(work), nbytes, _lib.stream_ptr(t.device)))
    return out


def irfft_shift(q, n=None):
    """``np.fft.fftshift(np.fft.irfft(q, axis=0), axes=0)`` -- the last line of the reference's examples."""
    return irfft(q, n=n, shift=True)


def stft(x, nfft, hop, window=None, bins=None):
    """Short-time Fourier transform ``X[f, q, t] = rfft(window * x[q, t*hop : t*hop + nfft])[f]`` of the
    microphone signals x [Q, S] (real), as a complex tensor [F, Q, T] with the frame axis innermost -- the
    layout `modal.steering.pwd` consumes.  `bins` keeps the first F bins (default nfft/2 + 1)."""
    lib = _lib.require_cuda()
    dev = _device_of(x)
    t = x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    if t.is_complex():
        raise ValueError("stft needs real signals")
    if t.ndim == 1:
        t = t[None, :]
    if t.ndim != 2:
        raise ValueError("x must have shape (Q, S)")
    is_double = t.dtype == torch.float64
    rdtype = torch.float64 if is_double else torch.float32
    t = t.to(device=dev, dtype=rdtype).contiguous()
    nfft, hop = int(nfft), int(hop)
    if nfft < 2 or nfft & (nfft - 1) or nfft > 16384:
        raise ValueError("nfft must be a power of two between 2 and 16384")
    Q, S = t.shape
    if S < nfft or hop < 1:
        raise ValueError("need at least one full frame (S >= nfft) and hop >= 1")
    T = 1 + (S - nfft) // hop
    F = nfft // 2 + 1 if bins is None else int(bins)
    if not 1 <= F <= nfft // 2 + 1:
        raise ValueError("bins must be in 1 .. nfft/2 + 1")
    win = None
    if window is not None:
        win = (window.detach() if isinstance(window, torch.Tensor) else torch.as_tensor(np.asarray(window)))
        win = win.to(device=dev, dtype=rdtype).contiguous()
        if win.shape != (nfft,):
            raise ValueError("window must have nfft entries")
    X = torch.empty((F, Q, T), dtype=torch.complex128 if is_double else torch.complex64, device=dev)
    nbytes = lib.sfa_stft_work_bytes(nfft, int(is_double))
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_stft(Q, S, nfft, hop, F, _lib.ptr(t), _lib.ptr(win), _lib.ptr(X), int(is_double),
                                _lib.ptr(work), nbytes, _lib.stream_ptr(dev)))
    return X
