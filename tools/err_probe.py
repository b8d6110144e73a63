import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import sfa_numpy_b200 as micarray
from oracle import sfa_oracle as O
from sfa_numpy_b200 import _lib
lib = _lib.load()
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
def bin_err(q, ref):
    q = q.reshape(q.shape[0], -1); ref = ref.reshape(ref.shape[0], -1)
    return float(np.max(np.max(np.abs(q - ref), axis=1) / np.max(np.abs(ref), axis=1)))
for (N, Q, L, F, T, r, setup) in ((32, 1089, 16384, 64, 32, 0.2, "rigid"), (16, 578, 2562, 64, 64, 0.1, "card")):
    az, co, w = O.fibonacci_grid(Q); azl, col, _ = O.fibonacci_grid(L)
    Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
    k = O.wavenumbers(4096)[::4096 // F][:F] + 0.1
    _, dn, _ = R.radial_filters(N, k, r, setup, 100.0, "Tikh")
    X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda")
    sel = [1, F // 2, F - 1]
    ref = O.pwd_factored(Y_q.cpu().numpy(), O.repeat_n_m(dn.cpu().numpy())[sel], Y_p.cpu().numpy(), X[sel].cpu().numpy().astype(np.complex128))
    for prec in ("c64", "c64_fast"):
        for env in ("", "1"):
            lib.sfa_set_option(b"no_kloop", 1.0 if env else 0.0)
            q = S.pwd(Y_q, dn, Y_p, X, precision=prec, form="factored")
            print("N=%d %s %s: bin_err %.3e" % (N, prec, "splitK" if env else "kloop", bin_err(q[sel].cpu().numpy(), ref)))
# random-data GEMM error vs K
rng = np.random.default_rng(0)
for K in (64, 128, 289, 578, 1089, 2048):
    B, Nn, Mm = 4, 256, 32
    Pn = (rng.standard_normal((1, Nn, K)) + 1j * rng.standard_normal((1, Nn, K))).astype(np.complex64)
    Rn = (rng.standard_normal((B, K, Mm)) + 1j * rng.standard_normal((B, K, Mm))).astype(np.complex64)
    P, Rt = torch.from_numpy(Pn).cuda(), torch.from_numpy(Rn).cuda()
    out = torch.zeros((B, Nn, Mm), dtype=torch.complex64, device="cuda")
    ref = np.matmul(Pn.astype(np.complex128), Rn.astype(np.complex128))
    res = []
    for hybrid in (0, 1):
        for env in ("", "1"):
            lib.sfa_set_option(b"no_kloop", 1.0 if env else 0.0)
            _lib.check(lib.sfa_cgemm3x(B, Nn, Mm, K, _lib.ptr(P), K, 0, _lib.ptr(Rt), Mm, K * Mm, _lib.ptr(out), Mm, Nn * Mm, None, Nn, hybrid, _lib.stream_ptr()))
            e = out.cpu().numpy() - ref
            res.append("%s/%s max %.2e rms %.2e" % ("hyb" if hybrid else "3x", "splitK" if env else "kloop", np.max(np.abs(e)) / np.max(np.abs(ref)), np.sqrt(np.mean(np.abs(e)**2) / np.mean(np.abs(ref)**2))))
    ref32 = np.matmul(Pn, Rn)
    e = ref32 - ref
    print("K=%d: " % K + " | ".join(res) + " | numpy c64 max %.2e rms %.2e" % (np.max(np.abs(e)) / np.max(np.abs(ref)), np.sqrt(np.mean(np.abs(e)**2) / np.mean(np.abs(ref)**2))))
