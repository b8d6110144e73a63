"""cfg 5 as one problem over the ranks with the complex beams gathered (dist.OverlappedPwd): NCCL against
peer-to-peer DMA pushes, with 0 / 8 / 16 SMs left free by the persistent kernels.  torchrun, 2+ ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray                    # noqa: E402
from sfa_numpy_b200 import _lib, dist as sd          # noqa: E402
from oracle import sfa_oracle as O                   # noqa: E402  (grids only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
A, R, S = micarray.modal.angular, micarray.modal.radial, micarray.modal.steering
N, Qn, F, L, T = 32, 1089, 4096, 16384, 32
az, co, w = O.fibonacci_grid(Qn)
azl, col, _ = O.fibonacci_grid(L)
Y_p, Y_q = A.sht_matrix(N, az, co, w), A.sht_matrix(N, azl, col)
_, dn, _ = R.radial_filters(N, O.wavenumbers(F) + 0.1, 0.2, "rigid", 100.0, "Tikh")
gen = torch.Generator(device=dev).manual_seed(5)
X = torch.randn((F, Qn, T), dtype=torch.complex64, device=dev, generator=gen)
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def timed(fn, n=3):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


f0, f1 = sd.shard_bins(F, world, rank)
plan = S.PwdPlan(Y_q, dn[f0:f1], Y_p, T, precision="c64_fast")
out = S.empty_beams(f1 - f0, L, T, torch.complex64, dev)
Xs = X[f0:f1].contiguous()
ms = timed(lambda: plan(Xs, out=out))
if rank == 0:
    print("resident slab, no exchange: %.2f ms" % ms, flush=True)
del plan, out
for transport in ("nccl", "p2p"):
    op = sd.OverlappedPwd(Y_q, dn, Y_p, T, pieces=pieces, precision="c64_fast", device=dev, transport=transport)
    Xbc = X[op.my_bins()].contiguous()
    for reserve in (0, 8, 16):
        lib.sfa_set_option(b"reserve_sms", float(reserve))
        ms = timed(lambda: op(Xbc))
        ing = 8.0 * F * L * op.buf.shape[2] * (world - 1) / world
        if rank == 0:
            print("%-5s pieces %d reserve_sms %2d: %.2f ms  (ingest %.0f GB/s per GPU)" % (transport, pieces, reserve, ms, ing / ms / 1e6), flush=True)
    lib.sfa_set_option(b"reserve_sms", 0.0)
    del op, Xbc
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
