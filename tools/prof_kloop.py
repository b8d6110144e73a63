"""Small driver for ncu: the factored-form contraction of the cfg 5 shape (N=32, 1089 mics, 16384 look
directions, T=32) on fewer bins -- two launches of cgemm3x_kloop_kernel (Z = d * Y_Q^H X, q = Y_L Z)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
PREC = sys.argv[2] if len(sys.argv) > 2 else "c64_fast"
N, Q, L, T = 32, 1089, 16384, 32
az, co, w = O.fibonacci_grid(Q)
azl, col, _ = O.fibonacci_grid(L)
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, O.wavenumbers(4096)[:F] + 0.1, 0.2, "rigid", 100.0, "Tikh")
X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda")
plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=PREC)
out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
for _ in range(2):
    plan(X, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan(X, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flops = 8.0 * F * T * M * (Q + L) if (M := (N + 1) ** 2) else 0
print(PREC, "F=%d: %.3f ms -> %.3e outputs/s, %.1f algorithmic TFLOP/s" % (F, ms, F * L * T / ms * 1e3, flops / ms * 1e-9))
