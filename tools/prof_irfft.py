"""ncu driver: one irfft_shift of the headline shape (F = 1024 -> n = 2048) on a column slab."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfa_numpy_b200 import fft as sfft   # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 960
q = torch.randn((1024, C), dtype=torch.complex64, device="cuda")
for _ in range(2):
    out = sfft.irfft(q, n=2048, shift=True)
torch.cuda.synchronize()
print(out.shape)
