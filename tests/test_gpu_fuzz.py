"""Seeded fuzz of the two tcgen05 GEMM kernels through the C ABI: random batch counts, ragged row / column /
inner extents (resident-K kernel, split-K passes, in-kernel K loop with one or several accumulation blocks,
single CTAs and CTA pairs, folded batches), against a float64 NumPy product.  The persistent kernels hand
work between warps through mbarrier phases, so shape-dependent task counts are what this is after."""
import numpy as np
import pytest

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
