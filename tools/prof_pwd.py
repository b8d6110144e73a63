"""Small driver for ncu: a few chunks of the headline steering contraction (cfg 3 shape, fewer bins)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 74
PREC = sys.argv[2] if len(sys.argv) > 2 else "c64"
N, L, T = 8, 2562, 938
azi, colat, w = O.grid_equal_angle(3)
azl, col, _ = O.fibonacci_grid(L)
k = O.wavenumbers(1024)[:F] + 0.5
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
Y_p = A.sht_matrix(N, azi, colat, w)
Y_q = A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
X = torch.from_numpy(O.synth_spectra(F, 64, T, seed=0)).cuda()
plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=PREC)
out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
for _ in range(3):
    plan(X, out=out)
torch.cuda.synchronize()
import ctypes
from sfa_numpy_b200 import _lib
lib = _lib.load()
lib.sfa_prof_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan(X, out=out)
e1.record()
torch.cuda.synchronize()
tot, cnt = ctypes.c_double(), ctypes.c_longlong()
lib.sfa_prof_read(ctypes.byref(tot), ctypes.byref(cnt))
span = ctypes.c_double()
lib.sfa_prof_span(ctypes.byref(span))
if os.environ.get("SFA_TIMELINE"):
    buf = (ctypes.c_double * (3 * 256))()
    n = lib.sfa_prof_timeline(buf, 256)
    rows = sorted((buf[3 * i + 1], buf[3 * i + 2], int(buf[3 * i])) for i in range(n))
    for a, b, k in rows:
        print("    %-5s %8.3f -> %8.3f ms  (%.3f)" % ("gemm" if k == 0 else "build", a, b, b - a))
lib.sfa_prof_enable(0)
print("  gemm kernels: %d launches, %.3f ms total, first start -> last end %.3f ms" % (cnt.value, tot.value, span.value))
print(PREC, "F=%d: %.3f ms -> %.3e outputs/s" % (F, e0.elapsed_time(e1), F * L * T / e0.elapsed_time(e1) * 1e3))
