"""K-loop kernel: time and error of cfg 4 / cfg 5 (stage 3, factored form) against the accumulation-block length."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray            # noqa: E402
from sfa_numpy_b200 import _lib              # noqa: E402
from oracle import sfa_oracle as O           # noqa: E402

lib = _lib.load()
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering


def bin_err(q, ref):
    q = q.reshape(q.shape[0], -1)
    ref = ref.reshape(ref.shape[0], -1)
    return float(np.max(np.max(np.abs(q - ref), axis=1) / np.max(np.abs(ref), axis=1)))


for (name, N, grid, L, F, T, r, setup) in (("cfg4", 16, "gauss", 2562, 1024, 64, 0.1, "card"), ("cfg5", 32, 1089, 16384, 4096, 32, 0.2, "rigid")):
    az, co, w = O.grid_gauss(N) if grid == "gauss" else O.fibonacci_grid(grid)
    azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, 100.0, "Tikh")
    Q = Y_p.shape[0]
    X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda")
    sel = [1, F // 2, F - 1]
    ref = O.pwd_factored(Y_q.cpu().numpy(), O.repeat_n_m(dn.cpu().numpy())[sel], Y_p.cpu().numpy(),
                         X[sel].cpu().numpy().astype(np.complex128))
    out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
    for prec in ("c64_fast", "c64"):
        for blk in ((0, 9, 12, 14, 18) if prec == "c64_fast" else (0, 6, 8, 10)):
            lib.sfa_set_option(b"kloop_block", float(blk))
            plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=prec)
            for _ in range(2):
                plan(X, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                plan(X, out=out)
            e1.record()
            torch.cuda.synchronize()
            print("%s %-8s kloop_block %2d: %.3f ms  err %.2e" % (name, prec, blk, e0.elapsed_time(e1) / 3,
                                                               bin_err(out[sel].cpu().numpy(), ref)), flush=True)
            del plan
    lib.sfa_set_option(b"kloop_block", 0.0)
    del X, out
    torch.cuda.empty_cache()
