"""GPU parity of the time <-> frequency ends: `fft.irfft_shift` against the time-domain output
of the four unmodified example scripts (golden `examples_time.npz`, np.fft.fftshift(np.fft.irfft(q, axis=0),
axes=0), doc/examples/modal_beamforming_rigid_spherical_array.py:40) and against np.fft on random spectra;
`fft.stft` against the NumPy restatement.  Tolerances: 1e-10 (complex128) / 1e-5 (complex64), relative to
the largest output sample."""
import os

import numpy as np
import pytest

from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cpu(t):
    return t.detach().cpu().numpy()


def relmax(x, ref):
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300))


@pytest.mark.parametrize("name", ["open_circular", "rigid_circular", "open_spherical", "rigid_spherical"])
def test_examples_time_domain(golden, name):
    import sfa_numpy_b200 as micarray
    gt = np.load(os.path.join(GOLDEN, "examples_time.npz"))
    q = golden["examples"][name + "_q"]
    ref = gt[name + "_q_t"]
    out = micarray.fft.irfft_shift(q)                       # complex128 -> float64, n = 2 (100 - 1) = 198 (Bluestein)
    assert tuple(out.shape) == ref.shape and cpu(out).dtype == np.float64
    assert relmax(cpu(out), ref) < 1e-10
    out32 = micarray.fft.irfft_shift(q.astype(np.complex64))
    assert cpu(out32).dtype == np.float32
    assert relmax(cpu(out32), ref) < 1e-5


@pytest.mark.parametrize("F,C,n", [(1025, 37, None), (1024, 10, 2048), (33, 1, None), (9, 5, 15), (9, 5, 16), (9, 4, 7),
                                   (100, 3, 198), (100, 6, 256), (300, 2, 100), (2, 3, None), (513, 64, 1024),
                                   (4097, 3, None), (70, 9, 139), (1024, 5, None), (2049, 4, None)])
def test_irfft_random(F, C, n):
    import sfa_numpy_b200 as micarray
    from sfa_numpy_b200 import _lib
    rng = np.random.default_rng(F * 131 + C)
    q = rng.standard_normal((F, C)) + 1j * rng.standard_normal((F, C))
    n_eff = 2 * (F - 1) if n is None else n
    for shift in (False, True):
        ref = np.fft.irfft(q, n=n, axis=0)
        if shift:
            ref = np.fft.fftshift(ref, axes=0)
        if n_eff > 4096:
            # documented limit of the FP64 kernel (shared-memory FFT): power-of-two length <= 4096
            with pytest.raises(_lib.SfaError):
                micarray.fft.irfft(q, n=n, shift=shift)
        else:
            out = micarray.fft.irfft(q, n=n, shift=shift)
            assert tuple(out.shape) == ref.shape
            assert relmax(cpu(out), ref) < 1e-10, (F, C, n, shift)
        out32 = micarray.fft.irfft(q.astype(np.complex64), n=n, shift=shift)
        assert relmax(cpu(out32), ref) < 1e-5, (F, C, n, shift)


def test_irfft_beam_tensor_and_errors():
    """[F, L, T] beam tensors (trailing axes flattened), row-padded views, the NumPy error for n < 1."""
    import torch
    import sfa_numpy_b200 as micarray
    S = micarray.modal.steering
    rng = np.random.default_rng(5)
    F, L, T = 65, 7, 138
    q = (rng.standard_normal((F, L, T)) + 1j * rng.standard_normal((F, L, T))).astype(np.complex64)
    buf = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))       # row-padded view (pitch 160)
    buf.copy_(torch.from_numpy(q))
    out = micarray.fft.irfft_shift(buf)
    ref = O.irfft_shift(q.astype(np.complex128))
    assert tuple(out.shape) == (128, L, T) and relmax(cpu(out), ref) < 1e-5
    with pytest.raises(ValueError):
        micarray.fft.irfft(np.ones((1, 4), dtype=np.complex128))              # n = 2 (1 - 1) = 0: NumPy raises too
    with pytest.raises(ValueError):
        np.fft.irfft(np.ones((1, 4), dtype=np.complex128), axis=0)
    assert tuple(micarray.fft.irfft(np.ones((5, 0), dtype=np.complex64)).shape) == (8, 0)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_stft(dtype, tol):
    import sfa_numpy_b200 as micarray
    rng = np.random.default_rng(2)
    for (Q, S, nfft, hop, bins) in [(3, 2000, 256, 64, None), (1, 512, 512, 512, None), (5, 5000, 1024, 300, 400),
                                    (2, 9000, 2048, 512, 1024), (4, 700, 64, 16, None)]:
        x = rng.standard_normal((Q, S)).astype(dtype)
        win = np.hanning(nfft).astype(dtype)
        for w in (None, win):
            ref = O.stft(x.astype(np.float64), nfft, hop, None if w is None else w.astype(np.float64), bins)
            X = micarray.fft.stft(x, nfft, hop, w, bins)
            assert tuple(X.shape) == ref.shape
            assert relmax(cpu(X), ref) < tol, (Q, S, nfft, hop, bins, w is None)


def test_stft_pwd_irfft_chain():
    """The whole 10-s-audio pipeline shape on a small case: time signals -> STFT -> PWD -> one frame back to the
    time domain, every stage on the device, against the NumPy/oracle chain."""
    import torch
    import sfa_numpy_b200 as micarray
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    N, nfft, hop = 3, 128, 32
    az, co, w = O.grid_gauss(N)
    Q = len(az)
    azl = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((Q, 1000))
    k = 2 * np.pi * np.arange(nfft // 2 + 1) * 48000.0 / nfft / 343.0
    X = micarray.fft.stft(x, nfft, hop)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, np.pi / 2)
    _, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "softclip")
    q = S.pwd(Y_q, dn, Y_p, X, precision="c128")
    q_t = micarray.fft.irfft_shift(q)
    Xr = O.stft(x, nfft, hop)
    dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100.0, "softclip")
    qr = O.pwd_factored(O.sht_matrix(N, azl, np.pi / 2), O.repeat_n_m(dn_ref), O.sht_matrix(N, az, co, w), Xr)
    ref = O.irfft_shift(qr)
    assert tuple(q_t.shape) == ref.shape == (nfft, 40, Xr.shape[2])
    assert relmax(cpu(q_t), ref) < 1e-9
