"""irfft / stft timing at the headline sizes: radix-16 fast path against the generic kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib, fft as sfft  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
HBM = 6561.6


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for F, n, C in ((1024, 2048, 256 * 938), (1025, 2048, 256 * 960), (513, 1024, 512 * 960), (2049, 4096, 128 * 960), (129, 256, 4096 * 960)):
    q = torch.randn((F, C), dtype=torch.complex64, device=dev)
    for generic in (0.0, 1.0):
        lib.sfa_set_option(b"fft_generic", generic)
        ms = t(lambda: sfft.irfft(q, n=n, shift=True))
        nbytes = 8.0 * F * C + 4.0 * n * C
        print("irfft F=%d n=%d C=%d %-8s %.3f ms  %.0f GB/s  %.2f of HBM peak" % (
            F, n, C, "generic" if generic else "radix-16", ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / HBM), flush=True)
    lib.sfa_set_option(b"fft_generic", 0.0)
    del q
T = 938
S_len = 2048 + 512 * (T - 1)
x = torch.randn((64, S_len), device=dev)
win = torch.hann_window(2048, device=dev)
for generic in (0.0, 1.0):
    lib.sfa_set_option(b"fft_generic", generic)
    ms = t(lambda: sfft.stft(x, 2048, 512, window=win, bins=1024))
    nbytes = 4.0 * 64 * S_len + 8.0 * 1024 * 64 * T
    print("stft 64 x %d nfft 2048 hop 512 %-8s %.3f ms  %.0f GB/s" % (S_len, "generic" if generic else "radix-16", ms, nbytes / ms / 1e6))
lib.sfa_set_option(b"fft_generic", 0.0)
