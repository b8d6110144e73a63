"""Timing probe for an fp16-split variant of the fused kernel (debug build): hi*hi pass at half the k-steps."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
from sfa_numpy_b200 import _lib
from oracle import sfa_oracle as O
F, N, L, T = 1024, 8, 2562, 938
azi, colat, w = O.grid_equal_angle(3); azl, col, _ = O.fibonacci_grid(L)
k = O.wavenumbers(1024)[:F] + 0.5
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
Y_p, Y_q = A.sht_matrix(N, azi, colat, w), A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
X = torch.from_numpy(O.synth_spectra(F, 64, T, seed=0)).cuda()
plan = S.PwdPlan(Y_q, dn, Y_p, T, precision="c64_fast")
out = S.empty_beams(F, L, T, torch.complex64, torch.device("cuda"))
pm = torch.empty((F, L), dtype=torch.float32, device="cuda")
lib = _lib.load()
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    for name, dbg in (("hybrid tf32 + bf16 (as shipped)", 0), ("16-bit hi*hi pass, half the k-steps (probe)", 256)):
        lib.sfa_set_option(b"fused_debug", float(dbg))
        a = t(lambda: plan(X, out=out))
        b = t(lambda: plan(X, power=pm, beams=False))
        print("%-48s full %.3f ms   map-only %.3f ms" % (name, a, b), flush=True)
lib.sfa_set_option(b"fused_debug", 0.0)
