import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import sfa_numpy_b200 as micarray
from oracle import sfa_oracle as O
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
N, F, L = 8, 512, 2562
azi, colat, w = O.grid_equal_angle(3); azl, col, _ = O.fibonacci_grid(L)
Yq = A.sht_matrix(N, azl, col).cpu().numpy(); Yp = A.sht_matrix(N, azi, colat, w).cpu().numpy()
dn = R.radial_filters(N, O.wavenumbers(F) + 0.1, 0.042, "rigid", 100.0, "Tikh")[1].cpu().numpy()
for T in (938, 960):
    host = S.PwdHost(Yq, dn, Yp, T, precision="c64_fast", device=0, chunk_bins=0)
    Xp = host.pinned_empty((F, 64, T)); Xp[...] = 1.0
    outp = host.pinned_empty((F, L, T))
    host(Xp, out=outp)
    t0 = time.perf_counter()
    for _ in range(2): host(Xp, out=outp)
    dt = (time.perf_counter() - t0) / 2
    print("T=%d: %.1f ms, D2H %.1f GB/s, %.3e outputs/s" % (T, dt * 1e3, outp.nbytes / dt / 1e9, F * L * T / dt))
    host.close()
