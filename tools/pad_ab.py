"""Same-process A/B of the fused steering kernel's row ends: whole 128-byte lines (out_pad_writable, the row padding
completed with zeros) against rows clipped at frame T, interleaved repetitions; with a debug build
(SFA_B200_LIB=.../libsfa_b200_dbg.so) also the store stream alone (no MMAs, no operand loads, no build) and the
kernel without stores."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
DBG = "dbg" in os.environ.get("SFA_B200_LIB", "")
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
masks = [(0, "full kernel")] + ([(26, "stores only"), (8, "no MMAs"), (2, "no build"), (1, "no stores")] if DBG else [])
variants = [(m, mn, pad, noleg) for (m, mn) in masks for pad in ((True,) if DBG else (True, False)) for noleg in (0, 1)]
res = {v: [] for v in variants}
for rep in range(REPS):
    for v in variants:
        m, mn, pad, noleg = v
        if DBG:
            lib.sfa_set_option(b"fused_debug", float(m))
        lib.sfa_set_option(b"no_legendre", float(noleg))
        plan._prepared_ld = None                      # the build path is chosen by the preparation
        plan(X, out=out, pad_writable=pad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            plan(X, out=out, pad_writable=pad)
        e1.record()
        torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / 4)
if DBG:
    lib.sfa_set_option(b"fused_debug", 0.0)
for v in variants:
    r = sorted(res[v])
    print("%-14s %-22s %-22s median %.3f ms  min %.3f  max %.3f" % (
        v[1], "whole-line row ends" if v[2] else "rows clipped at T", "group-product table" if v[3] else "Legendre build",
        r[len(r) // 2], r[0], r[-1]))
lib.sfa_set_option(b"no_legendre", 0.0)
