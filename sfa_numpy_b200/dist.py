"""Frequency-bin sharding across the GPUs of one box (one process per GPU, torch.distributed).

Bins are independent in stage 3 (and in stage 2 apart from the zero-replacement search, which is
why stage 2 is simply replicated: it is a few MB), so rank r owns the contiguous slab
``[f0, f1)`` of the F bins, computes ``q[f0:f1]`` with no communication, and the only collective
is an optional all-gather of the output slabs (NCCL over NVLink on GPUs; gloo in the CPU tests).
Stage 1 (SH matrices) is recomputed on every rank -- microseconds, cheaper than a broadcast.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_bins(F, world_size, rank):
    """Contiguous, balanced slab [f0, f1) of F bins for `rank` (first F % world ranks get one more)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(F), int(world_size))
    f0 = rank * base + min(rank, extra)
    return f0, f0 + base + (1 if rank < extra else 0)


def slab_sizes(F, world_size):
    return [shard_bins(F, world_size, r)[1] - shard_bins(F, world_size, r)[0] for r in range(world_size)]


def _default_compute(Y_look, dn, Y_mic, X, **kw):
    from .modal import steering
    return steering.pwd(Y_look, dn, Y_mic, X, **kw)


def pwd_sharded(Y_look, dn, Y_mic, X, gather=False, group=None, compute=None, **kw):
    """Plane-wave decomposition with the bins sharded over the ranks of `group`.

    X and dn may be the full [F, ...] arrays (each rank slices its slab) -- pass
    ``local=True`` in kw-free form by giving arrays whose first axis already equals the slab.
    Returns the local slab [f1-f0, L, T]; with ``gather=True`` returns the full [F, L, T] on
    every rank (one all-gather of the slabs, the only collective on the path).
    """
    compute = compute or _default_compute
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    F = dn.shape[0]
    f0, f1 = shard_bins(F, world, rank)
    X_loc = X[f0:f1] if X.shape[0] == F else X
    if X_loc.shape[0] != f1 - f0:
        raise ValueError("X must hold all {} bins or exactly this rank's {} bins".format(F, f1 - f0))
    q_loc = compute(Y_look, dn[f0:f1], Y_mic, X_loc, **kw)
    if not gather or world == 1:
        return q_loc
    return all_gather_slabs(q_loc, F, group=group)


def all_gather_slabs(q_loc, F, group=None):
    """All-gather of contiguous [F_r, L, T] slabs into [F, L, T] (uneven slabs are padded to the
    largest slab for the collective and trimmed afterwards)."""
    world = dist.get_world_size(group)
    sizes = slab_sizes(F, world)
    big = max(sizes)
    tail = tuple(q_loc.shape[1:])
    if all(s == big for s in sizes):
        out = torch.empty((F,) + tail, dtype=q_loc.dtype, device=q_loc.device)
        dist.all_gather_into_tensor(_as_real(out), _as_real(q_loc.contiguous()), group=group)
        return out
    padded = torch.zeros((big,) + tail, dtype=q_loc.dtype, device=q_loc.device)
    padded[: q_loc.shape[0]] = q_loc
    buf = torch.empty((world * big,) + tail, dtype=q_loc.dtype, device=q_loc.device)
    dist.all_gather_into_tensor(_as_real(buf), _as_real(padded), group=group)
    return torch.cat([buf[r * big: r * big + sizes[r]] for r in range(world)], dim=0)


class OverlappedPwd:
    """Strong-scaled plane-wave decomposition of ONE [F]-bin problem with the complex beams gathered on every
    rank, the collective hidden behind the compute.

    The bins are dealt out block-cyclically: piece c of rank r is the bin block
    ``[(c * world + r) * cb, ... + cb)`` with ``cb = F / (world * pieces)``.  Rank r computes its block of piece c
    straight into its place in the full ``[F, L, T]`` tensor; as soon as that kernel has finished (event on the
    compute stream) the in-place ``all_gather_into_tensor`` of piece c -- one contiguous ``[world * cb, L, ld]``
    slab -- runs on a communication stream while piece c + 1 is being computed.  Every GPU has to ingest
    (world - 1) / world of the output over NVLink, so for the HBM-bound configs the gather, not the compute,
    sets the time; the overlap hides the smaller of the two.

    Two transports for the exchange:
      "nccl" -- in-place ``all_gather_into_tensor`` per piece on a communication stream.  NCCL's copy kernels need
                SMs, and the persistent tensor-core kernels own all of them, so the gather of piece c mostly runs
                in the gaps between kernels (measured at 2 GPUs, cfg 5: 50 ms against 37 ms of compute);
      "p2p"  -- every rank maps every peer's output buffer (CUDA IPC through torch's tensor-sharing handles) and
                PUSHES its finished block into each peer with DMA copies (copy engines over NVLink /
                NVSwitch, one stream per peer).  No SM is involved, so the transfer of piece c really runs under
                the kernels of piece c + 1.  A one-element all-reduce at both ends of a call orders the pushes
                against the peers' use of their buffers.  Every block leaves the GPU N - 1 times, though: at 8 GPUs
                the pushes reach 400 GB/s per GPU where NCCL (one send through the switch) reaches 560-650;
      "auto" -- "p2p" for two ranks on GPUs, "nccl" otherwise.

    `compute(plan, X_piece, out_piece)` may be injected (the CPU/gloo test of the scheduling logic does)."""

    def __init__(self, Y_look, dn, Y_mic, T, pieces=4, precision="c64_fast", form="auto", group=None, make_plan=None,
                 device=None, transport="nccl"):
        if transport not in ("nccl", "p2p", "auto"):
            raise ValueError("transport must be 'nccl', 'p2p' or 'auto'")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if transport == "auto":
            # measured on 8 x B200 (cfg 5): pushes win while a rank has one or two peers (2 GPUs: 38.9 against
            # 50.1 ms); from four GPUs on every block leaves a rank N - 1 times while NCCL sends it once through
            # the switch (8 GPUs: 26.7 ms with NCCL against 38.0 ms)
            on_gpu = (device is not None and torch.device(device).type == "cuda") or (device is None and dn.is_cuda)
            transport = "p2p" if (on_gpu and self.world == 2) else "nccl"
        self.transport = transport
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        F = dn.shape[0]
        if F % (self.world * pieces):
            raise ValueError("F = {} must be a multiple of world * pieces = {}".format(F, self.world * pieces))
        self.pieces, self.cb, self.F, self.T = int(pieces), F // (self.world * pieces), F, int(T)
        self.dn = dn
        L = Y_look.shape[0]
        if make_plan is None:
            from .modal import steering
            self.plan = steering.PwdPlan(Y_look, dn[: self.cb], Y_mic, T, precision=precision, form=form, device=device)
            dev, cdtype, ld = self.plan.dev, self.plan.cdtype, steering._padded_pitch(int(T))
        else:
            self.plan, dev, cdtype, ld = make_plan(Y_look, dn[: self.cb], Y_mic, T)
        # `buf` is this object's own row-padded allocation: the kernels may complete the last line of every row
        self._plan_kw = {"pad_writable": True} if make_plan is None else {}
        self.buf = torch.empty((F, L, ld), dtype=cdtype, device=dev)          # row-padded; gathered with its padding
        self.full = self.buf[:, :, : self.T]
        self.comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.peers = None
        if transport == "p2p" and self.world > 1:
            if dev.type != "cuda":
                raise ValueError("the p2p transport needs CUDA devices")
            self.peers = map_peer_buffers(self.buf, group=group)
            self.push_streams = [torch.cuda.Stream(device=dev) if r != self.rank else None for r in range(self.world)]
            self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
            self._lib = _lib.require_cuda()

    def block_start(self, piece, rank=None):
        return (piece * self.world + (self.rank if rank is None else rank)) * self.cb

    def my_bins(self):
        """Global bin indices of this rank, in the order its X must be laid out ([pieces * cb, Q, T])."""
        return [self.block_start(c) + i for c in range(self.pieces) for i in range(self.cb)]

    def _call_p2p(self, X_loc):
        cb, W = self.cb, self.world
        cur = torch.cuda.current_stream()
        dist.all_reduce(self._flag, group=self.group)          # every rank has entered the call: its buffer may be written
        gate = torch.cuda.Event()
        gate.record()
        for s in self.push_streams:
            if s is not None:
                s.wait_event(gate)
        for c in range(self.pieces):
            f0 = self.block_start(c)
            self.plan.dn = self.dn[f0:f0 + cb]
            self.plan(X_loc[c * cb:(c + 1) * cb], out=self.full[f0:f0 + cb], **self._plan_kw)
            ev = torch.cuda.Event()
            ev.record()
            mine = self.buf[f0:f0 + cb]
            nbytes = mine.numel() * mine.element_size()
            off = mine.data_ptr() - self.buf.data_ptr()
            for k in range(1, W):                              # staggered: rank r starts with peer r + 1
                r = (self.rank + k) % W
                s = self.push_streams[r]
                s.wait_event(ev)
                # cudaMemcpyPeerAsync on this device's stream s: copy engines, no kernel
                _lib.check(self._lib.sfa_peer_copy(ctypes.c_void_p(self.peers[r].data_ptr() + off),
                                                   self.peers[r].device.index, ctypes.c_void_p(mine.data_ptr()),
                                                   self.buf.device.index, nbytes, ctypes.c_void_p(s.cuda_stream)))
        for s in self.push_streams:
            if s is not None:
                cur.wait_stream(s)
        dist.all_reduce(self._flag, group=self.group)          # every rank's pushes have landed
        return self.full

    def __call__(self, X_loc):
        if self.peers is not None:
            return self._call_p2p(X_loc)
        cb, W = self.cb, self.world
        for c in range(self.pieces):
            f0 = self.block_start(c)
            self.plan.dn = self.dn[f0:f0 + cb]
            self.plan(X_loc[c * cb:(c + 1) * cb], out=self.full[f0:f0 + cb], **self._plan_kw)
            if W == 1:
                continue
            slab = self.buf[c * W * cb:(c + 1) * W * cb]
            mine = self.buf[f0:f0 + cb]
            if self.comm is not None:
                ev = torch.cuda.Event()
                ev.record()
                self.comm.wait_event(ev)
                with torch.cuda.stream(self.comm):
                    dist.all_gather_into_tensor(_as_real(slab), _as_real(mine), group=self.group)
            else:                                                              # CPU tests (gloo): no in-place aliasing
                dist.all_gather_into_tensor(_as_real(slab), _as_real(mine.clone()), group=self.group)
        if self.comm is not None and W > 1:
            torch.cuda.current_stream().wait_stream(self.comm)
        return self.full


def map_peer_buffers(local, group=None):
    """Map the tensor `local` of every rank into every other rank's process (CUDA IPC, one box): returns a list
    with, at index r, a tensor that aliases rank r's buffer (the local tensor itself at this rank's index).
    The handles travel through ``all_gather_object``; torch's own tensor-sharing machinery
    (``torch.multiprocessing.reductions``) creates and opens them, so the caching allocator stays in charge of
    the memory.  The peer tensors live on the PEER's device index: copies into them are peer-to-peer DMA."""
    from torch.multiprocessing.reductions import reduce_tensor
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    fn, args = reduce_tensor(local)
    handles = [None] * world
    dist.all_gather_object(handles, (fn, args), group=group)
    peers, err = [], None
    try:
        for r in range(world):
            if r == rank:
                peers.append(local)
            else:
                f, a = handles[r]
                peers.append(f(*a))
    except Exception as e:                                     # e.g. no IPC / peer access between two of the devices
        err = e
    # all ranks agree on the outcome (a rank that raised alone would leave the others waiting in the next collective);
    # the reduction also keeps every buffer alive until all ranks have opened it
    ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=local.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        raise RuntimeError("peer mapping of the output buffers failed on at least one rank" +
                           (": %r" % (err,) if err is not None else ""))
    return peers


def gather_power_map(q_loc, F, group=None, reduce=None):
    """Frame-averaged beam power map mean_t |q|^2, [F, L] float32, gathered on every rank: the
    cheap collective when only the map (not the beam signals) is needed downstream.  The local
    reduction is `steering.power_map` (one CUDA kernel); `reduce` injects a replacement (the CPU
    tests of the sharding logic do)."""
    if reduce is None:
        from .modal import steering
        reduce = steering.power_map
    p_loc = reduce(q_loc)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return p_loc
    return all_gather_slabs(p_loc, F, group=group)


def _as_real(t):
    return torch.view_as_real(t) if t.is_complex() else t
