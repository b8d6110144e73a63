"""Two ranks, two GPUs, NCCL: the frequency-sharded plane-wave decomposition with the all-gather of the output
(dist.pwd_sharded(gather=True), dist.OverlappedPwd, dist.gather_power_map) against the single-GPU result.
Needs two devices: skipped on a one-GPU box (NCCL refuses two ranks on one device); run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sfa_numpy_b200 as micarray
        from sfa_numpy_b200 import dist as sd
        from oracle import sfa_oracle as O
        A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
        N, F, L, T = 4, 24, 300, 200
        az, co, w = O.grid_gauss(N)
        azl, col, _ = O.fibonacci_grid(L)
        k = O.wavenumbers(64)[:F] + 0.4
        Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
        _, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
        X = torch.from_numpy(O.synth_spectra(F, len(az), T, seed=3)).to(dev)
        single = S.pwd(Y_q, dn, Y_p, X, precision="c64_fast")                 # whole problem on this GPU
        # (a) contiguous slabs + one all-gather
        gathered = sd.pwd_sharded(Y_q, dn, Y_p, X, gather=True, precision="c64_fast")
        assert gathered.shape == single.shape and torch.equal(gathered, single)
        # (b) block-cyclic pieces, in-place all-gather overlapped with the compute
        op = sd.OverlappedPwd(Y_q, dn, Y_p, T, pieces=3, precision="c64_fast")
        got = op(X[op.my_bins()].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(got, single)
        # (b2) the same with peer-to-peer DMA pushes instead of NCCL (buffers mapped over CUDA IPC), twice (buffer reuse)
        op2 = sd.OverlappedPwd(Y_q, dn, Y_p, T, pieces=3, precision="c64_fast", transport="p2p")
        Xbc = X[op2.my_bins()].contiguous()
        for _ in range(2):
            op2.full.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            got2 = op2(Xbc)
            torch.cuda.synchronize()
            assert torch.equal(got2, single)
        # (c) the power map out of the fused kernel, gathered
        f0, f1 = sd.shard_bins(F, world, rank)
        plan = S.PwdPlan(Y_q, dn[f0:f1], Y_p, T, precision="c64_fast")
        plan(X[f0:f1].contiguous(), power=True)
        pm = sd.all_gather_slabs(plan.power, F)
        ref = (single.abs() ** 2).mean(dim=-1)
        assert float(((pm - ref).abs() / ref.amax(dim=1, keepdim=True)).max()) < 1e-5
        # against the oracle too (rank 0)
        if rank == 0:
            dn_ref, _ = O.regularize(1 / O.spherical_pw(N, k, 0.042, "rigid"), 100.0, "Tikh")
            ref64 = O.pwd_factored(O.sht_matrix(N, azl, col), O.repeat_n_m(dn_ref), O.sht_matrix(N, az, co, w),
                                   X.cpu().numpy().astype(np.complex128))
            err = np.max(np.abs(gathered.cpu().numpy() - ref64)) / np.max(np.abs(ref64))
            assert err < 1e-5
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_nccl_world2_sharded_pwd():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not run two ranks on one device)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert results.get(0) and results.get(1)


def test_two_devices_in_one_process():
    """One process driving two GPUs (function attributes such as the dynamic shared-memory opt-in are per
    device): every kernel family runs on cuda:1 after it has run on cuda:0, and the results agree bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import sfa_numpy_b200 as micarray
    from oracle import sfa_oracle as O
    A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
    res = []
    for dev in ("cuda:0", "cuda:1"):
        d = torch.device(dev)
        out = {}
        # fused steering kernel (> 48 KB dynamic shared memory), stream-ahead path included (many bins)
        N, Q, L, F, T = 3, 16, 389, 260, 100
        az, co, w = O.fibonacci_grid(Q)
        azl, col, _ = O.fibonacci_grid(L)
        Y_p = A.sht_matrix(N, torch.from_numpy(az).to(d), torch.from_numpy(co).to(d), torch.from_numpy(w).to(d))
        Y_q = A.sht_matrix(N, torch.from_numpy(azl).to(d), torch.from_numpy(col).to(d))
        k = torch.from_numpy(np.linspace(0.5, 40.0, F)).to(d)
        _, dn, _ = R.radial_filters(N, k, 0.042, "rigid", 100.0, "Tikh")
        X = torch.from_numpy(O.synth_spectra(F, Q, T, seed=2)).to(d)
        assert Y_p.device == d and dn.device == d
        out["fused"] = S.pwd(Y_q, dn, Y_p, X, precision="c64_fast", form="steering").cpu()
        # K-loop kernel (K > 64) and the resident-K kernel (factored form), FP64 path, cht rows kernel, FFT fast path
        N2, Q2 = 9, 100
        az2, co2, w2 = O.fibonacci_grid(Q2)
        Yp2 = A.sht_matrix(N2, torch.from_numpy(az2).to(d), torch.from_numpy(co2).to(d), torch.from_numpy(w2).to(d))
        Yq2 = A.sht_matrix(N2, torch.from_numpy(azl).to(d), torch.from_numpy(col).to(d))
        _, dn2, _ = R.radial_filters(N2, k[:8], 0.1, "rigid", 100.0, "Tikh")
        X2 = torch.from_numpy(O.synth_spectra(8, Q2, 40, seed=3)).to(d)
        out["kloop"] = S.pwd(Yq2, dn2, Yp2, X2, precision="c64", form="factored").cpu()
        out["c128"] = S.pwd(Yq2, dn2, Yp2, X2.to(torch.complex128), precision="c128").cpu()
        out["cht"] = A.cht_matrix(14, torch.linspace(0, 6.28, 5000, dtype=torch.float64, device=d)).cpu()
        out["irfft"] = micarray.fft.irfft(out["fused"][:129, :3, :].to(d).contiguous(), n=256, shift=True).cpu()
        torch.cuda.synchronize(d)
        res.append(out)
    for key in res[0]:
        assert torch.equal(res[0][key], res[1][key]), key
