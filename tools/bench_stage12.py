"""SH / Bessel evals per second (BASELINE.json metric, second half) for stages 1 and 2.

Each line: config, GPU time (CUDA events, best of n), evals/s, achieved store bandwidth vs the
measured HBM peak, and the oracle (SciPy, single thread) on the same or a bounded input."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_time(fn, n=6, inner=8):
    """Seconds per call: `inner` calls are enqueued back to back between two events, so the host-side
    cost of a call (~35 us of Python + ctypes) overlaps the previous call's kernel; for kernels shorter
    than that the figure is the call rate of the API, not the kernel duration."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best * 1e-3


def cpu_time(fn, n=2):
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def run(cpu=True, hbm_gbs=6561.6):
    import sfa_numpy_b200 as micarray
    from oracle import sfa_oracle as O
    A, R = micarray.modal.angular, micarray.modal.radial
    rows = []
    dev = torch.device("cuda")
    for name, N, P, Pcpu in (("cfg3 look dirs", 8, 2562, 2562), ("cfg4 look dirs", 16, 2562, 2562),
                             ("cfg5 look dirs", 32, 16384, 2048), ("roofline sweep", 32, 1 << 17, 0)):
        azi, colat, w = O.fibonacci_grid(P)
        d = [torch.from_numpy(a).to(dev) for a in (azi, colat, w)]
        M = (N + 1) ** 2
        t = gpu_time(lambda: A.sht_matrix(N, d[0], d[1], d[2]))
        row = {"stage": "sht_matrix", "config": "%s N=%d P=%d" % (name, N, P), "gpu_s": t, "evals_per_s": P * M / t,
               "GBps": (16.0 * P * M + 24.0 * P) / t / 1e9}
        row["frac_hbm"] = row["GBps"] / hbm_gbs
        if cpu and Pcpu:
            tc = cpu_time(lambda: O.sht_matrix(N, azi[:Pcpu], colat[:Pcpu], w[:Pcpu]), 1)
            row["cpu_evals_per_s"] = Pcpu * M / tc
        rows.append(row)
    pol = np.linspace(0, 2 * np.pi, 1 << 20, endpoint=False)
    dpol = torch.from_numpy(pol).to(dev)
    t = gpu_time(lambda: A.cht_matrix(16, dpol))
    rows.append({"stage": "cht_matrix", "config": "N=16 P=2^20", "gpu_s": t, "evals_per_s": (1 << 20) * 33 / t,
                 "GBps": 16.0 * (1 << 20) * 33 / t / 1e9, "frac_hbm": 16.0 * (1 << 20) * 33 / t / 1e9 / hbm_gbs})
    for name, N, F, setup, kind, Fcpu in (("cfg3", 8, 1024, "rigid", "spherical_pw", 1024),
                                          ("cfg5", 32, 4096, "rigid", "spherical_pw", 1024),
                                          ("cfg4 sweep 2^22", 16, 1 << 22, "card", "spherical_pw", 4096),
                                          ("cfg2", 16, 1024, "open", "circular_pw", 1024),
                                          ("circ sweep 2^20", 16, 1 << 20, "rigid", "circular_pw", 0)):
        kr = np.geomspace(0.01, 50, F)
        r = 0.1
        dk = torch.from_numpy(kr / r).to(dev)
        C = N + 1 if kind == "spherical_pw" else 2 * N + 1
        t = gpu_time(lambda: R.radial_filters(N, dk, r, setup, 100.0, "Tikh", kind=kind, check="defer"), n=5)
        R.check_deferred()
        row = {"stage": kind + "+regularize(fused)", "config": "%s N=%d F=%d %s" % (name, N, F, setup), "gpu_s": t,
               "evals_per_s": F * C / t, "GBps": (8.0 * F + 40.0 * F * C) / t / 1e9}
        row["frac_hbm"] = row["GBps"] / hbm_gbs
        if cpu and Fcpu:
            fn = O.spherical_pw if kind == "spherical_pw" else O.circular_pw
            tc = cpu_time(lambda: O.regularize(1 / fn(N, kr[:Fcpu] / r, r, setup), 100.0, "Tikh"), 1)
            row["cpu_evals_per_s"] = Fcpu * C / tc
        rows.append(row)
    return rows


if __name__ == "__main__":
    for r in run():
        print(json.dumps(r))
