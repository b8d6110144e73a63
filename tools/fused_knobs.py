"""Same-process A/B of library options on the fused steering kernel (cfg 3), interleaved repetitions.
usage: [SFA_B200_LIB=.../libsfa_b200_dbg.so] python tools/fused_knobs.py BINS REPS "name:key=val,key=val" ...
(a variant with no options: "name:")."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402  (input generator only)

F, REPS = int(sys.argv[1]), int(sys.argv[2])
variants = []
for spec in sys.argv[3:]:
    name, _, opts = spec.partition(":")
    variants.append((name, dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in opts.split(",") if kv)))
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
ref = None
res = {name: [] for name, _ in variants}
for rep in range(REPS):
    for name, opts in variants:
        for kk, v in opts.items():
            _lib.check(lib.sfa_set_option(kk.encode(), v))
        plan._prepared_ld = None
        plan(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            plan(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 4)
        if rep == 0 and not opts.get("fused_debug"):
            chk = out[::97, ::31].clone()
            if ref is None:
                ref = chk
            elif not torch.equal(chk, ref):
                print("NOTE %s: output differs from the first variant (max %.3e)" % (name, float((chk - ref).abs().max())))
        for kk in opts:
            lib.sfa_set_option(kk.encode(), 0.0)
for name, _ in variants:
    v = sorted(res[name])
    print("%-40s median %.3f ms  min %.3f  max %.3f" % (name, v[len(v) // 2], v[0], v[-1]))
