"""cfg 2 (circular array, 33 filter columns) through the fused kernel: rotation build against the table build
(and, with a debug library, against the kernel with the build switched off)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
from sfa_numpy_b200 import _lib
from oracle import sfa_oracle as O
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
N, Q, F, L, T = 16, 32, 1024, 360, 256
pol = np.linspace(0, 2 * np.pi, Q, endpoint=False)
Y_p, Y_q = A.cht_matrix(N, pol, np.full(Q, 1.0 / Q)), A.cht_matrix(N, np.linspace(0, 2 * np.pi, L, endpoint=False))
_, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, 0.1, "open", 3000.0, "softclip", kind="circular_pw")
X = torch.randn((F, Q, T), dtype=torch.complex64, device="cuda")
lib = _lib.load()
DBG = "dbg" in os.environ.get("SFA_B200_LIB", "")
variants = [("rotation build", {}), ("table build", {"no_legendre": 1}), ("prepass", {"fused_prepass": 1}),
            ("D=74 ring=222", {"fused_D": 74, "fused_ring": 222}), ("D=18 ring=56", {"fused_D": 18, "fused_ring": 56}),
            ("no cluster", {"no_cluster": 1})] + ([("no build", {"fused_debug": 2}), ("no MMAs", {"fused_debug": 8}), ("no stores", {"fused_debug": 1})] if DBG else [])
res = {}
for rep in range(4):
    for name, opts in variants:
        for k, v in opts.items():
            lib.sfa_set_option(k.encode(), float(v))
        plan = S.PwdPlan(Y_q, dn, Y_p, T, precision="c64_fast", form="steering")
        out = S.empty_beams(*plan.out_shape, torch.complex64, torch.device("cuda"))
        plan(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan(X, out=out)
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(name, []).append(e0.elapsed_time(e1) / 5)
        if rep == 0:
            res[name + " chk"] = out[::50, ::7].clone()
        for k in opts:
            lib.sfa_set_option(k.encode(), 0.0)
for name, _ in variants:
    v = sorted(res[name])
    print("%-16s median %.3f ms  min %.3f" % (name, v[len(v) // 2], v[0]))
print("rotation == table bitwise:", bool(torch.equal(res["rotation build chk"], res["table build chk"])),
      " max diff %.3e" % float((res["rotation build chk"] - res["table build chk"]).abs().max()))
