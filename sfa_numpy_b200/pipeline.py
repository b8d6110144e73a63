"""The whole hot path of the reference's example scripts as ONE CUDA-graph launch per batch of spectra.

What every example does (doc/examples/modal_beamforming_rigid_spherical_array.py:18-39):

    Y_p = sht_matrix(N, azi, colat, weights)            # stage 1, microphones
    Y_q = sht_matrix(N, azi_pwd, colat_pwd)             # stage 1, look directions
    bn  = spherical_pw(N, k, r, setup)                  # stage 2
    dn, _ = regularize(1 / bn, a0, method)
    q   = matmul(matmul(matmul(Y_q, diagonal_mode_mat(dn)), conj(Y_p.T)), p)     # stage 3

At the BASELINE sizes stages 1-2 are launch-latency-bound (a few microseconds of work behind ~10 launches),
so the sequence is captured once into a CUDA graph (all kernels are launched through the C ABI on torch's
capture stream; descriptors and workspaces are static) and replayed with a single launch.  A graph, not a
tracing compiler: the kernels are the hand-written ones, the graph only removes the launch gaps.
"""
import torch

from . import _lib
from .modal import angular, radial, steering
from .util import _device_of


class ModalBeamformer:
    """Fixed geometry, streaming spectra.

    mic : (azi, colat, weights) for spherical arrays, (pol, weights) for circular ones (weights may be None)
    look: (azi, colat) or (pol,)
    k   : wavenumbers of ALL F bins (stage 2 is replicated: the zero replacement searches across bins)
    bins: optional (f0, f1): the slab of bins this instance (this GPU) computes in stage 3
    power: also produce the beam power map mean_t |q|^2 [f1 - f0, L] (fused into the steering kernel)
    `X` / `out` / `power` are the static buffers of the graph: fill `X`, call `run()`, read `out`.
    """

    def __init__(self, N, mic, look, k, r, setup, a0, method, T, kind="spherical", precision="c64_fast", bins=None,
                 power=False, graph=True, device=None):
        dev = device if device is not None else _device_of(*[m for m in mic if m is not None], k)
        self.dev = dev
        self.N, self.r, self.setup, self.a0, self.method, self.kind = int(N), float(r), setup, float(a0), method, kind
        f64 = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64).to(dev).contiguous()  # noqa: E731
        self.mic = tuple(f64(a) for a in mic)
        self.look = tuple(f64(a) for a in look)
        self.k = f64(k).reshape(-1)
        F = self.k.numel()
        self.bins = (0, F) if bins is None else (int(bins[0]), int(bins[1]))
        if not 0 <= self.bins[0] <= self.bins[1] <= F:
            raise ValueError("bins must be a sub-range of the {} wavenumbers".format(F))
        self.T = int(T)
        self.precision = precision
        cdtype = torch.complex128 if precision == "c128" else torch.complex64
        # one eager pass of stages 1-2 with the reference's error behaviour (ValueError when a zero of the radial
        # function cannot be replaced, radial.py:209), which also sizes the static buffers
        Y_p, Y_q, dn = self._stage12(check="now")
        nb = self.bins[1] - self.bins[0]
        Q, L = Y_p.shape[0], Y_q.shape[0]
        self.X = torch.zeros((nb, Q, self.T), dtype=cdtype, device=dev)
        self.out = steering.empty_beams(nb, L, self.T, cdtype, dev)
        self.power = torch.zeros((nb, L), dtype=torch.float32, device=dev) if power else None
        self.plan = steering.PwdPlan(Y_q, dn[self.bins[0]:self.bins[1]], Y_p, self.T, precision=precision, device=dev)
        self.graph = None
        self.graph_error = None
        lib = _lib.load()
        n0 = lib.sfa_launch_count()
        self._eager()
        self.launches_per_run = int(lib.sfa_launch_count() - n0)     # kernels per step (eager or replayed)
        if graph:
            self._capture()

    def _stage12(self, check):
        if self.kind == "spherical":
            Y_p = angular.sht_matrix(self.N, *self.mic)
            Y_q = angular.sht_matrix(self.N, *self.look)
            rk = "spherical_pw"
        elif self.kind == "circular":
            Y_p = angular.cht_matrix(self.N, *self.mic)
            Y_q = angular.cht_matrix(self.N, *self.look)
            rk = "circular_pw"
        else:
            raise ValueError("kind must be 'spherical' or 'circular'")
        _, dn, _ = radial.radial_filters(self.N, self.k, self.r, self.setup, self.a0, self.method, kind=rk, check=check)
        if dn.ndim == 1:
            dn = dn[None, :]
        return Y_p, Y_q, dn

    def _eager(self):
        Y_p, Y_q, dn = self._stage12(check="none")
        p = self.plan
        p.Y_look, p.Y_mic, p.dn = Y_q, Y_p, dn[self.bins[0]:self.bins[1]]
        p(self.X, out=self.out, power=self.power)
        self.Y_mic, self.Y_look, self.dn = Y_p, Y_q, dn

    def _capture(self):
        try:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._eager()                             # warm-up on the capture-side stream
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager()
            self.graph = g
        except Exception as e:                            # keep working eagerly, but say so
            self.graph = None
            self.graph_error = repr(e)
            torch.cuda.synchronize(self.dev)

    def run(self):
        """One step over the spectra currently in `self.X`; returns `self.out` (asynchronous)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
        return self.out

    def run_eager(self):
        self._eager()
        return self.out

    def __call__(self, X):
        self.X.copy_(X)
        return self.run()
