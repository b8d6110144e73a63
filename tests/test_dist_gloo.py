"""CPU, world_size 2, gloo: the bin-sharding + all-gather host logic of sfa_numpy_b200.dist.
The per-slab compute is injected (a plain torch einsum) because the CUDA path needs a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfa_numpy_b200 import dist as sd


def test_shard_bins_partition():
    for F in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            edges = [sd.shard_bins(F, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sd.slab_sizes(F, world)
    with pytest.raises(ValueError):
        sd.shard_bins(8, 2, 2)


def _torch_pwd(Y_look, dn, Y_mic, X, **kw):
    # per-column diagonal form: q_f = Y_L diag(d_f) Y_Q^H X_f
    Z = torch.einsum("qm,fqt->fmt", Y_mic.conj(), X) * dn[:, :, None]
    return torch.einsum("lm,fmt->flt", Y_look, Z)


def _worker(rank, world, port, F, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        L, M, Q, T = 5, 4, 3, 6
        Y_look = torch.randn(L, M, dtype=torch.complex128, generator=g)
        Y_mic = torch.randn(Q, M, dtype=torch.complex128, generator=g)
        dn = torch.randn(F, M, dtype=torch.complex128, generator=g)
        X = torch.randn(F, Q, T, dtype=torch.complex128, generator=g)
        full = _torch_pwd(Y_look, dn, Y_mic, X)
        loc = sd.pwd_sharded(Y_look, dn, Y_mic, X, compute=_torch_pwd)
        f0, f1 = sd.shard_bins(F, world, rank)
        assert loc.shape[0] == f1 - f0 and torch.allclose(loc, full[f0:f1])
        # slab passed pre-sliced
        loc2 = sd.pwd_sharded(Y_look, dn, Y_mic, X[f0:f1], compute=_torch_pwd)
        assert torch.equal(loc, loc2)
        gathered = sd.pwd_sharded(Y_look, dn, Y_mic, X, gather=True, compute=_torch_pwd)
        assert gathered.shape == full.shape and torch.allclose(gathered, full)
        pm = sd.gather_power_map(loc.to(torch.complex64), F,
                                 reduce=lambda q: (q.real.float() ** 2 + q.imag.float() ** 2).mean(dim=-1))
        with pytest.raises(ValueError):                  # the product reduction is CUDA-only: no CPU fallback
            sd.gather_power_map(loc.to(torch.complex64), F)
        assert pm.shape == (F, L)
        assert torch.allclose(pm, (full.abs() ** 2).mean(-1).float(), rtol=1e-4)
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("F", [8, 7])
def test_gloo_world2(F):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, F, results), nprocs=2, join=True)
    assert results.get(0) and results.get(1)


class _TorchPlan:
    """Stand-in for steering.PwdPlan in the CPU test of OverlappedPwd's scheduling (block-cyclic pieces,
    in-place slab all-gather): same call signature, plain torch arithmetic."""

    def __init__(self, Y_look, dn, Y_mic, T):
        self.Y_look, self.Y_mic, self.dn = Y_look, Y_mic, dn

    def __call__(self, X, out=None):
        out.copy_(_torch_pwd(self.Y_look, self.dn, self.Y_mic, X))
        return out


def _overlap_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        F, L, M, Q, T, pieces = 12, 5, 4, 3, 6, 3
        Y_look = torch.randn(L, M, dtype=torch.complex128, generator=g)
        Y_mic = torch.randn(Q, M, dtype=torch.complex128, generator=g)
        dn = torch.randn(F, M, dtype=torch.complex128, generator=g)
        X = torch.randn(F, Q, T, dtype=torch.complex128, generator=g)
        full = _torch_pwd(Y_look, dn, Y_mic, X)
        op = sd.OverlappedPwd(Y_look, dn, Y_mic, T, pieces=pieces,
                              make_plan=lambda yl, d, ym, t: (_TorchPlan(yl, d, ym, t), torch.device("cpu"),
                                                              torch.complex128, t + 2))
        bins = op.my_bins()
        assert len(bins) == F // world and len(set(bins)) == len(bins)
        got = op(X[bins])
        assert got.shape == full.shape and torch.allclose(got, full)
        # every bin is owned by exactly one rank
        owners = [None] * world
        dist.all_gather_object(owners, bins)
        assert sorted(b for o in owners for b in o) == list(range(F))
        with pytest.raises(ValueError):
            sd.OverlappedPwd(Y_look, dn[:10], Y_mic, T, pieces=pieces, make_plan=lambda *a: None)
        with pytest.raises(ValueError):                      # peer-to-peer pushes need CUDA devices
            sd.OverlappedPwd(Y_look, dn, Y_mic, T, pieces=pieces, transport="p2p",
                             make_plan=lambda yl, d, ym, t: (_TorchPlan(yl, d, ym, t), torch.device("cpu"),
                                                             torch.complex128, t + 2))
        with pytest.raises(ValueError):
            sd.OverlappedPwd(Y_look, dn, Y_mic, T, pieces=pieces, transport="carrier pigeon", make_plan=lambda *a: None)
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_overlapped_pwd_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_overlap_worker, args=(2, port, results), nprocs=2, join=True)
    assert results.get(0) and results.get(1)
