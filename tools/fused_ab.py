"""Same-process A/B of the fused steering kernel's spectra conversion: pre-pass for all bins against stream-ahead
conversion with different look-ahead distances / ring sizes, interleaved repetitions (the part-to-part and
run-to-run spread of this kernel is larger than the effects)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
variants = [("pre-pass", dict(fused_prepass=1)), ("stream-ahead default", {}),
            ("D=8 ring=24", dict(fused_D=8, fused_ring=24)), ("D=16 ring=24", dict(fused_D=16, fused_ring=24)),
            ("D=16 ring=32", dict(fused_D=16, fused_ring=32)), ("D=16 ring=64", dict(fused_D=16, fused_ring=64)),
            ("D=24 ring=48", dict(fused_D=24, fused_ring=48))]
res = {name: [] for name, _ in variants}
for rep in range(REPS):
    for name, opts in variants:
        for kk, v in opts.items():
            lib.sfa_set_option(kk.encode(), float(v))
        plan(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            plan(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 4)
        for kk in opts:
            lib.sfa_set_option(kk.encode(), 0.0)
for name, _ in variants:
    v = sorted(res[name])
    print("%-24s median %.3f ms  min %.3f  max %.3f   (whole stage-3 call, %d bins)" % (name, v[len(v) // 2], v[0], v[-1], F))
