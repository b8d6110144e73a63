"""Peer-copy bandwidth probe (2+ ranks under torchrun): how fast does a rank push a 1 GiB slab into a peer's
IPC-mapped buffer (a) with torch's cross-device copy_, (b) with cudaMemcpyPeerAsync through sfa_peer_copy,
(c) with NCCL all-gather for reference."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfa_numpy_b200 import _lib, dist as sd          # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
n = 1 << 30
buf = torch.empty(world * n, dtype=torch.uint8, device=dev)
peers = sd.map_peer_buffers(buf)
mine = buf[rank * n:(rank + 1) * n]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    dist.barrier()
    return dt


def push_torch():
    for k in range(1, world):
        r = (rank + k) % world
        peers[r][rank * n:(rank + 1) * n].copy_(mine, non_blocking=True)


streams = [torch.cuda.Stream(device=dev) for _ in range(world)]


def push_abi():
    for k in range(1, world):
        r = (rank + k) % world
        dst = peers[r].data_ptr() + rank * n
        _lib.check(lib.sfa_peer_copy(ctypes.c_void_p(dst), peers[r].device.index, ctypes.c_void_p(mine.data_ptr()), local,
                                     n, ctypes.c_void_p(streams[r].cuda_stream)))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)


def nccl():
    dist.all_gather_into_tensor(buf, mine)


for name, fn in (("torch copy_", push_torch), ("sfa_peer_copy", push_abi), ("nccl all_gather", nccl)):
    try:
        dt = timed(fn)
        print("rank %d  %-16s %7.2f ms  egress %.1f GB/s" % (rank, name, dt * 1e3, (world - 1) * n / dt / 1e9), flush=True)
    except Exception as e:
        print("rank %d  %-16s failed: %r" % (rank, name, e), flush=True)
dist.barrier()
dist.destroy_process_group()
