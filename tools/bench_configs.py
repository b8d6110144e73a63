"""Stage-3 time of every BASELINE config at full size (beam outputs/s), both tensor-core modes."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import sfa_numpy_b200 as micarray
from oracle import sfa_oracle as O
from test_gpu_fullsize import CONFIGS
CONFIGS = dict(CONFIGS, cfg3=("sph", 8, "ea3", 1024, 2562, 938, 0.042, "rigid", 100.0, "Tikh"))
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None     # e.g. cfg1,cfg2 (the steering form of cfg 5 takes 40 s)
for name in sorted(CONFIGS):
    if ONLY and name not in ONLY:
        continue
    kind, N, grid, F, L, T, r, setup, a0, method = CONFIGS[name]
    if kind == "sph":
        az, co, w = (O.grid_equal_angle(N) if grid == "equal_angle" else O.grid_equal_angle(3) if grid == "ea3"
                     else O.grid_gauss(N) if grid == "gauss" else O.fibonacci_grid(grid))
        azl, col, _ = O.fibonacci_grid(L)
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, a0, method)
    else:
        pol = np.linspace(0, 2 * np.pi, grid, endpoint=False)
        Y_p, Y_q = A.cht_matrix(N, pol, np.full(grid, 1.0 / grid)), A.cht_matrix(N, np.linspace(0, 2 * np.pi, L, endpoint=False))
        _, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, r, setup, a0, method, kind="circular_pw")
    Q = Y_p.shape[0]
    X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda")
    forms = sys.argv[1].split(",") if len(sys.argv) > 1 else ["auto"]
    for prec, form in [(p_, f_) for p_ in ("c64", "c64_fast") for f_ in forms]:
        try:
            plan = S.PwdPlan(Y_q, dn, Y_p, T, precision=prec, form=form)
        except Exception as e:          # e.g. steering form with more than 64 groups
            print(json.dumps({"config": name, "precision": prec, "form": form, "error": str(e)[:80]}))
            continue
        out = S.empty_beams(*plan.out_shape, torch.complex64, torch.device("cuda"))
        for _ in range(2):
            plan(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            plan(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"config": name, "N": N, "Q": Q, "F": F, "L": L, "T": T, "precision": prec, "form": form, "form_used": plan.shape.form,
                          "ms": round(ms, 3), "outputs_per_s": F * L * T / ms * 1e3}))
    del X, out, plan
    torch.cuda.empty_cache()
