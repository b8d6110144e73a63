"""Seeded fuzz of the two tcgen05 GEMM kernels through the C ABI: random batch counts, ragged row / column /
inner extents (resident-K kernel, split-K passes, in-kernel K loop with one or several accumulation blocks,
single CTAs and CTA pairs, folded batches), against a float64 NumPy product.  The persistent kernels hand
work between warps through mbarrier phases, so shape-dependent task counts are what this is after."""
import numpy as np
import pytest

from oracle import sfa_oracle as O

pytestmark = pytest.mark.gpu


def test_cgemm3x_fuzz():
    import torch
    from sfa_numpy_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(20261020)
    worst = 0.0
    for case in range(48):
        B = int(rng.integers(1, 10))
        Nn = int(rng.integers(1, 420))
        Mm = int(rng.choice([1, 7, 32, 33, 64, 100, 128, 130, 300]))
        K = int(rng.choice([1, 9, 31, 64, 65, 96, 100, 257, 288, 289, 321, 700]))
        shared = bool(rng.integers(0, 2)) or K > 257          # keep the split-K (per-batch P) cases small
        hybrid = int(rng.integers(0, 2))
        use_scale = bool(rng.integers(0, 2))
        Pn = (rng.standard_normal((1 if shared else B, Nn, K)) + 1j * rng.standard_normal((1 if shared else B, Nn, K))).astype(np.complex64)
        Rn = (rng.standard_normal((B, K, Mm)) + 1j * rng.standard_normal((B, K, Mm))).astype(np.complex64)
        sc = (rng.standard_normal((B, Nn)) + 1j * rng.standard_normal((B, Nn))).astype(np.complex64)
        P, R, S = (torch.from_numpy(a).cuda() for a in (Pn, Rn, sc))
        out = torch.full((B, Nn, Mm), float("nan"), dtype=torch.complex64, device="cuda")
        _lib.check(lib.sfa_cgemm3x(B, Nn, Mm, K, _lib.ptr(P), K, 0 if shared else Nn * K, _lib.ptr(R), Mm, K * Mm,
                                   _lib.ptr(out), Mm, Nn * Mm, _lib.ptr(S) if use_scale else None, Nn, hybrid,
                                   _lib.stream_ptr()))
        ref = np.matmul(Pn.astype(np.complex128), Rn.astype(np.complex128))
        if use_scale:
            ref = ref * sc[:, :, None].astype(np.complex128)
        got = out.cpu().numpy()
        assert np.isfinite(got.view(np.float32)).all(), (case, B, Nn, Mm, K, shared, hybrid)
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        assert err < 6e-6, "case %d %s shared=%s hybrid=%d scale=%s: %.3e" % (case, (B, Nn, Mm, K), shared, hybrid, use_scale, err)
    print("fuzz worst normwise error %.2e" % worst)


def test_fused_kernel_fuzz():
    """Seeded random shapes through the fused steering kernel with enough bins for the stream-ahead conversion
    (plane ring, hand-over counters, tail rows): bit-equal to the pre-pass variant, spot bins against the oracle,
    repeated launches identical (a race in the hand-over would show up as a mismatch between any of these)."""
    import torch
    import sfa_numpy_b200 as micarray
    from sfa_numpy_b200 import _lib
    S = micarray.modal.steering
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    for case in range(10):
        N = int(rng.integers(1, 8))
        Q = int(rng.integers((N + 1) ** 2 // 2 + 1, 65))
        L = int(rng.choice([int(rng.integers(129, 400)), int(rng.integers(900, 1500)), 128 * int(rng.integers(2, 9)) + int(rng.integers(0, 17))]))
        T = int(rng.integers(1, 60)) * 2
        ltiles = (L - L % 128) // 128 if 0 < L % 128 <= 16 else -(-L // 128)
        lt_groups = (ltiles + 1) // 2 if ltiles >= 2 else 1
        in_flight = -(-(74 if ltiles >= 2 else 148) // lt_groups)
        F = 4 * in_flight + int(rng.integers(1, 40))
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Yp, Yq = O.sht_matrix(N, az, co, w), O.sht_matrix(N, azl, col)
        k = np.linspace(0.5, 40.0, F)
        dn, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100, "Tikh")
        X = torch.from_numpy((rng.standard_normal((F, Q, T)) + 1j * rng.standard_normal((F, Q, T))).astype(np.complex64)).cuda()
        prec = "c64_fast" if case % 2 == 0 else "c64"
        outs = {}
        for prepass in (0.0, 1.0):
            try:
                _lib.check(lib.sfa_set_option(b"fused_prepass", prepass))
                plan = S.PwdPlan(Yq, dn, Yp, T, precision=prec, form="steering")
                q1 = plan(X, power=True).clone()
                q2 = plan(X, power=True)
                assert torch.equal(q1, q2), ("repeat", case, N, Q, L, F, T, prepass)
                outs[prepass] = q1
            finally:
                lib.sfa_set_option(b"fused_prepass", 0.0)
        assert torch.equal(outs[0.0], outs[1.0]), ("stream-ahead vs pre-pass", case, N, Q, L, F, T)
        sel = sorted(set([0, F // 2, F - 1]))
        ref = O.pwd_factored(Yq, O.repeat_n_m(dn)[sel], Yp, X[sel].cpu().numpy().astype(np.complex128))
        got = outs[0.0][sel].cpu().numpy()
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        assert err < 1e-5, (case, N, Q, L, F, T, err)
