"""Timing attribution of the fused steering kernel (debug build: SFA_B200_LIB=.../libsfa_b200_dbg.so).
Runs the cfg-3 shape on F bins with parts of the kernel switched off (sfa_set_option("fused_debug", mask))."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 296
PREC = sys.argv[2] if len(sys.argv) > 2 else "c64_fast"
N, L, T = 8, 2562, 938
azi, colat, w = O.grid_equal_angle(3)
azl, col, _ = O.fibonacci_grid(L)
k = O.wavenumbers(1024)[:F] + 0.5
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
Y_p, Y_q = A.sht_matrix(N, azi, colat, w), A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
X = torch.from_numpy(O.synth_spectra(F, 64, T, seed=0)).cuda()
plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=PREC)
out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
lib = _lib.load()
NAMES = {1: "no TMA stores", 2: "no build FMAs", 4: "no staging writes", 8: "no MMAs", 16: "no operand loads"}
masks = [0, 1, 2, 4, 8, 16, 1 | 4, 8 | 16, 1 | 4 | 2, 1 | 2 | 4 | 16, 2 | 8 | 16, 31]
for m in masks:
    lib.sfa_set_option(b"fused_debug", float(m))
    for _ in range(2):
        plan(X, out=out)
    torch.cuda.synchronize()
    lib.sfa_prof_enable(1)
    for _ in range(3):
        plan(X, out=out)
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.sfa_prof_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.sfa_prof_enable(0)
    what = ", ".join(v for b, v in NAMES.items() if m & b) or "full kernel"
    print("mask %2d  %-60s %.3f ms / launch (%d bins)  -> %.2f ms per 1024 bins" % (
        m, what, tot.value / cnt.value, F, tot.value / cnt.value * 1024 / F))
lib.sfa_set_option(b"fused_debug", 0.0)
