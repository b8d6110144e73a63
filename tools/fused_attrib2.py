"""Timing attribution of the stream-ahead conversion of the fused steering kernel (debug build).
usage: SFA_B200_LIB=.../libsfa_b200_dbg.so python tools/fused_attrib2.py [bins]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N, L, T = 8, 2562, 938
azi, colat, w = O.grid_equal_angle(3)
azl, col, _ = O.fibonacci_grid(L)
k = O.wavenumbers(1024)[:F] + 0.5
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
Y_p, Y_q = A.sht_matrix(N, azi, colat, w), A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
X = torch.from_numpy(O.synth_spectra(F, 64, T, seed=0)).cuda()
plan = S.PwdPlan(Y_q, dn, Y_p, T, precision="c64_fast")
out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
lib = _lib.load()


def run(label, **opts):
    for kk, v in opts.items():
        lib.sfa_set_option(kk.encode(), float(v))
    for _ in range(2):
        plan(X, out=out)
    torch.cuda.synchronize()
    lib.sfa_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan(X, out=out)
    e1.record()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.sfa_prof_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.sfa_prof_enable(0)
    print("%-58s fused kernel %.3f ms   whole call %.3f ms" % (label, tot.value / cnt.value, e0.elapsed_time(e1) / 3))
    for kk in opts:
        lib.sfa_set_option(kk.encode(), 0.0)


for rep in range(2):
    run("pre-pass for all bins", fused_prepass=1)
    run("stream-ahead (default D, ring)")
    run("stream-ahead, converters read no X", fused_debug=32)
    run("stream-ahead, converters write no planes", fused_debug=64)
    run("stream-ahead, no X reads, no plane writes", fused_debug=96)
    run("stream-ahead, D=8", fused_D=8)
    run("stream-ahead, D=32 ring=96", fused_D=32, fused_ring=96)
    run("stream-ahead, D=16 ring=1024 (no reuse)", fused_D=16, fused_ring=1024)
    run("stream-ahead, no MMAs", fused_debug=8)
    run("pre-pass, no MMAs", fused_debug=8, fused_prepass=1)
    run("stream-ahead, no TMA stores", fused_debug=1)
    run("pre-pass, no TMA stores", fused_debug=1, fused_prepass=1)
