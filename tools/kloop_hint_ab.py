"""K-loop kernel (cfg 4 / cfg 5, factored form): output stores with and without the L2 evict_first policy,
interleaved repetitions; the results must be bit-equal."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402

lib = _lib.load()
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
for (name, N, grid, L, F, T, r, setup) in (("cfg4", 16, "gauss", 2562, 1024, 64, 0.1, "card"), ("cfg5", 32, 1089, 16384, 4096, 32, 0.2, "rigid")):
    az, co, w = O.grid_gauss(N) if grid == "gauss" else O.fibonacci_grid(grid)
    azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, 100.0, "Tikh")
    X = torch.randn((F, Y_p.shape[0], T), dtype=torch.complex64, device="cuda")
    out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
    plan = S.PwdPlan(Y_q, dn, Y_p, T, precision="c64_fast")
    res, chk = {0: [], 1: []}, {}
    for rep in range(4):
        for nohint in (0, 1):
            lib.sfa_set_option(b"kloop_no_hint", float(nohint))
            plan(X, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                plan(X, out=out)
            e1.record()
            torch.cuda.synchronize()
            res[nohint].append(e0.elapsed_time(e1) / 3)
            if rep == 0:
                chk[nohint] = out[::37, ::101].clone()
    lib.sfa_set_option(b"kloop_no_hint", 0.0)
    for nohint in (0, 1):
        v = sorted(res[nohint])
        print("%s %-22s median %.3f ms  min %.3f  max %.3f" % (name, "plain stores" if nohint else "evict_first stores",
                                                                v[len(v) // 2], v[0], v[-1]), flush=True)
    print(name, "bit-equal:", bool(torch.equal(chk[0], chk[1])))
    del X, out, plan
    torch.cuda.empty_cache()
