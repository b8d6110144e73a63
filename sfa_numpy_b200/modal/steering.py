"""Plane-wave decomposition / steering: ``q_f = Y_L diag(d_f) Y_Q^H X_f`` for every bin.

The reference has no function for this step -- every example script spells it
as a dense diagonal stack and three ``np.matmul`` calls
(doc/examples/modal_beamforming_rigid_spherical_array.py:36-39):

    D = diagonal_mode_mat(dn)
    A = np.matmul(np.matmul(Y_q, D), np.conj(Y_p.T))
    q = np.matmul(A, p[:, :, None])

``pwd(Y_q, dn, Y_p, p)`` computes exactly that ``q`` (same argument roles)
without ever forming ``D`` ([F, M, M]) or, at large sizes, ``A`` ([F, L, Q]).
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..util import as_complex_tensor, _device_of

PRECISIONS = {"c64": 0, "c128": 1, "c64_fast": 2}
FORMS = {"auto": 0, "steering": 1, "factored": 2}


def _shape_of(Y_look, dn, Y_mic, F, Q, T, precision, form):
    L, M = Y_look.shape
    if Y_mic.shape != (Q, M):
        raise ValueError("Y_mic must have shape (Q, M) = ({}, {}), got {}".format(Q, M, tuple(Y_mic.shape)))
    if dn.shape[0] != F:
        raise ValueError("dn must have one row per frequency bin ({}), got {}".format(F, dn.shape[0]))
    Cd = dn.shape[1]
    if Cd == M:
        per_order = 0
    elif Cd * Cd == M:
        per_order = 1                      # radial.diagonal_mode_mat semantics (repeat_n_m, radial.py:284)
    else:
        raise ValueError("dn has {} columns; expected M={} (per column) or sqrt(M) (per order)".format(Cd, M))
    return _lib.PwdShape(F, L, Q, T, M, Cd, per_order, PRECISIONS[precision], FORMS[form], 0)


_C64_DN_LIMIT = 1e30


def _check_c64_range(dn, precision):
    """The complex64 paths narrow the radial filters to float32.  Un-regularised filters (method 'none', or a
    huge a0) reach |1/b_n| ~ 1e47 at high orders and small kr, which is Inf in float32
    and would turn into NaN beams silently; the reference computes that case in complex128.  One tiny
    reduction + host read per plan / one-shot call -- not on the per-step path."""
    if precision == "c128" or dn.numel() == 0:
        return
    a = torch.view_as_real(dn).abs()
    if not bool(torch.isfinite(a).all()) or float(a.max()) > _C64_DN_LIMIT:
        raise ValueError("radial filters exceed the float32 range of the complex64 paths (max |dn| > 1e30 or "
                         "non-finite): regularise them (radial.regularize) or use precision='c128'")


def _padded_pitch(T):
    """Row pitch (in frames) of the beam tensor: multiples of 32 frames (256 B) keep every 128-byte
    line of a row tile inside one CTA, worth ~1.4x of HBM write bandwidth (stage 3)."""
    return T if T <= 128 or T % 32 == 0 else (T + 31) // 32 * 32


_PAD_ATTR = "_sfa_row_pad_owned"


def _alloc_out(F, L, T, cdtype, dev):
    ld = _padded_pitch(T)
    buf = torch.empty((F, L, ld), dtype=cdtype, device=dev)
    if ld == T:
        return buf
    # the padding frames [T, ld) of every row belong to this library: the kernels may complete the last 128-byte
    # line of a row with zeros (sfa_pwd_shape.out_pad_writable), which HBM absorbs 15-20 % faster than a partial line
    view = buf[:, :, :T]
    setattr(view, _PAD_ATTR, True)
    return view


def _pad_owned(out):
    """True for the row-padded views made by `empty_beams` / `pwd` themselves (a slice of one is a new tensor
    object and does not carry the mark: the kernels then leave everything beyond frame T alone)."""
    return bool(getattr(out, _PAD_ATTR, False))


def empty_beams(F, L, T, dtype=torch.complex64, device=None):
    """Uninitialised (F, L, T) beam tensor with the row pitch the kernels like best (a view of a
    row-padded buffer when T > 128 is not a multiple of 32); pass it as ``out=``."""
    return _alloc_out(F, L, T, dtype, device if device is not None else _device_of())


def _check_out(out, F, L, T, cdtype, dev):
    """`out` may be contiguous or a row-padded view: stride(2) == 1, stride(0) == L * stride(1)."""
    if (tuple(out.shape) != (F, L, T) or out.dtype != cdtype or out.device != dev or
            (T > 1 and out.stride(2) != 1) or out.stride(1) < T or out.stride(0) != L * out.stride(1)):
        raise ValueError("out must be a {} tensor of shape {} with unit frame stride and stride(0) == L*stride(1)"
                         .format(cdtype, (F, L, T)))
    return out.stride(1)


def power_map(q):
    """Frame-averaged beam power ``mean_t |q[f, l, t]|^2`` as an [F, L] float32 tensor -- what the
    reference's examples plot (``db(q)`` over look direction and frequency,
    doc/examples/modal_beamforming_rigid_spherical_array.py:43-57) and the small collective of the
    multi-GPU path.  `q` is a complex64 CUDA beam tensor as returned by `pwd` (row-padded views are
    fine); one HBM-read-bound kernel, no intermediate tensors."""
    if not (torch.is_tensor(q) and q.is_cuda and q.dtype == torch.complex64 and q.dim() == 3):
        raise ValueError("power_map needs a complex64 CUDA tensor of shape (F, L, T)")
    F, L, T = q.shape
    ld = _check_out(q, F, L, T, torch.complex64, q.device)
    out = torch.empty((F, L), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.sfa_power_map(F * L, T, ld, _lib.ptr(q), _lib.ptr(out), _lib.stream_ptr(q.device)))
    return out


def pwd(Y_look, dn, Y_mic, p, precision=None, form="auto", out=None):
    """Apply the modal beamformer to microphone spectra.

    Parameters
    ----------
    Y_look : (L, M) SH/CH matrix at the look directions (``Y_q`` in the examples).
    dn : (F, N+1) radial filters per order (spherical, expanded like ``diagonal_mode_mat``),
         or (F, M) per column (circular: ``circ_diagonal_mode_mat``).
    Y_mic : (Q, M) SH/CH matrix at the microphones, already carrying the quadrature
         weights (``Y_p`` in the examples); it enters conjugate-transposed.
    p : (F, Q) or (F, Q, T) microphone spectra.
    precision : "c64" (tensor cores, 3xTF32, FP32 accumulate; error ~1e-6 normwise),
         "c64_fast" (tensor cores, TF32 for hi*hi + one BF16 pass for the hi*lo cross terms: 2/3 of
         the tensor work, error still ~1e-6 normwise) or "c128" (FP64 on the CUDA cores).
         Default: from ``p``'s dtype (complex64 -> "c64", else "c128").
    Returns
    -------
    q : (F, L, T) beam outputs (``np.matmul(A, p[:, :, None])`` keeps the T axis, so T=1 for 2-D p).
        For T > 128 not a multiple of 32 the tensor is a view of a row-padded buffer (see
        `_padded_pitch`); ``q.contiguous()`` / ``q.cpu().numpy()`` give the dense array.
    """
    lib = _lib.require_cuda()
    dev = _device_of(p, Y_look, dn, Y_mic)
    if precision is None:
        pd = p.dtype if isinstance(p, torch.Tensor) else np.asarray(p).dtype
        precision = "c64" if pd in (torch.complex64, np.complex64) else "c128"
    if precision not in PRECISIONS:
        raise ValueError("precision must be 'c64', 'c64_fast' or 'c128'")
    if form not in FORMS:
        raise ValueError("form must be 'auto', 'steering' or 'factored'")
    cdtype = torch.complex128 if precision == "c128" else torch.complex64
    Y_look = as_complex_tensor(Y_look, dev)
    Y_mic = as_complex_tensor(Y_mic, dev)
    dn = as_complex_tensor(dn, dev)
    if dn.ndim == 1:
        dn = dn[None, :]
    X = as_complex_tensor(p, dev, cdtype)
    if X.ndim == 2:
        X = X[:, :, None]
    if X.ndim != 3:
        raise ValueError("p must have shape (F, Q) or (F, Q, T)")
    F, Q, T = X.shape
    shape = _shape_of(Y_look, dn, Y_mic, F, Q, T, precision, form)
    _check_c64_range(dn, precision)
    L = Y_look.shape[0]
    if out is None:
        out = _alloc_out(F, L, T, cdtype, dev)
    if F > 0 and L > 0:
        shape.out_ld = _check_out(out, F, L, T, cdtype, dev)
        shape.out_pad_writable = 1 if _pad_owned(out) else 0
    nbytes = lib.sfa_pwd_work_bytes(ctypes.byref(shape))
    if nbytes == 0:
        raise ValueError("unsupported plane-wave-decomposition shape")
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sfa_pwd_apply(ctypes.byref(shape), _lib.ptr(Y_look), _lib.ptr(Y_mic), _lib.ptr(dn),
                                     _lib.ptr(X), _lib.ptr(out), _lib.ptr(work), nbytes, _lib.stream_ptr(dev)))
    return out


def synthesize(Y_mic, bn, Y_src, precision="c128"):
    """Forward model of the examples: ``p = matmul(matmul(Y_p, D), conj(Y_pw.T))`` with
    ``D = diagonal_mode_mat(bn)`` (doc/examples/modal_beamforming_rigid_spherical_array.py:21-25) --
    the pressure at the microphones for unit sources at the directions of ``Y_src``.
    Returns (F, Q, S).  Same kernels as `pwd` (the sources play the role of the microphones and the
    "spectra" are the identity)."""
    dev = _device_of(Y_mic, bn, Y_src)
    Y_src = as_complex_tensor(Y_src, dev)
    bn_t = as_complex_tensor(bn, dev)
    F = 1 if bn_t.ndim == 1 else bn_t.shape[0]
    S_ = Y_src.shape[0]
    cdtype = torch.complex128 if precision == "c128" else torch.complex64
    eye = torch.eye(S_, dtype=cdtype, device=dev).expand(F, S_, S_).contiguous()
    return pwd(Y_mic, bn_t, Y_src, eye, precision=precision)


class PwdPlan:
    """Reusable device-side plan: workspace allocated once, operators kept resident.

    The operator preparation (group products of Y_look / Y_mic, `sfa_pwd_prepare`) runs when the plan is
    created and again only after `Y_look` / `Y_mic` have been re-assigned; `dn` may be re-assigned freely
    (it is read by every call).  ``plan(X, out=..., power=...)``: `power` is an optional float32 [F, L]
    tensor (or True to allocate one, then available as ``plan.power``) that receives the beam power map
    mean_t |q|^2 -- accumulated inside the fused steering kernel, i.e. without another pass over q."""

    def __init__(self, Y_look, dn, Y_mic, T, precision="c64", form="auto", device=None):
        self.lib = _lib.require_cuda()
        dev = device if device is not None else _device_of(Y_look, dn, Y_mic)
        self.dev = dev
        self.precision = precision
        self.cdtype = torch.complex128 if precision == "c128" else torch.complex64
        self._Y_look = as_complex_tensor(Y_look, dev)
        self._Y_mic = as_complex_tensor(Y_mic, dev)
        dn = as_complex_tensor(dn, dev)
        self.dn = dn[None, :] if dn.ndim == 1 else dn
        F, Q = self.dn.shape[0], self._Y_mic.shape[0]
        self.shape = _shape_of(self._Y_look, self.dn, self._Y_mic, F, Q, int(T), precision, form)
        _check_c64_range(self.dn, precision)      # checked once per plan; re-assigned `dn` is the caller's business
        self.nbytes = self.lib.sfa_pwd_work_bytes(ctypes.byref(self.shape))
        if self.nbytes == 0:
            raise ValueError("unsupported plane-wave-decomposition shape")
        self.work = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        self.out_shape = (F, self._Y_look.shape[0], int(T))
        self.power = None
        self._prepared_ld = None

    # re-assigning an operator invalidates the prepared group products
    @property
    def Y_look(self):
        return self._Y_look

    @Y_look.setter
    def Y_look(self, v):
        self._Y_look = as_complex_tensor(v, self.dev)
        self._prepared_ld = None

    @property
    def Y_mic(self):
        return self._Y_mic

    @Y_mic.setter
    def Y_mic(self, v):
        self._Y_mic = as_complex_tensor(v, self.dev)
        self._prepared_ld = None

    def prepare(self):
        """(Re)compute everything that depends on Y_look / Y_mic only."""
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.sfa_pwd_prepare(ctypes.byref(self.shape), _lib.ptr(self._Y_look), _lib.ptr(self._Y_mic),
                                                _lib.ptr(self.work), self.nbytes, _lib.stream_ptr(self.dev)))
        self._prepared_ld = self.shape.out_ld

    def __call__(self, X, out=None, power=None, beams=True, pad_writable=None):
        """``pad_writable``: whether the kernels may overwrite the row padding of `out` (frames [T, stride(1)) of
        every row) -- default: only for the row-padded tensors of `empty_beams`; pass True for a slice of one.
        ``beams=False`` (with `power`): only the beam power map is produced -- the fused steering kernel skips its
        stores, the 8 F L T bytes of beams never reach memory; returns the map.  Needs a plan that runs the fused
        kernel (complex64 precisions, steering form, at most 64 microphones), otherwise ValueError."""
        if X.dtype != self.cdtype or not X.is_contiguous() or X.device != self.dev:
            raise ValueError("X must be a contiguous {} tensor on {}".format(self.cdtype, self.dev))
        if tuple(X.shape) != (self.out_shape[0], self._Y_mic.shape[0], self.out_shape[2]):
            raise ValueError("X must have shape (F, Q, T) = {}".format((self.out_shape[0], self._Y_mic.shape[0],
                                                                         self.out_shape[2])))
        if not beams:
            if power is None or power is False:
                raise ValueError("beams=False needs power=True or a power tensor")
            if out is not None:
                raise ValueError("beams=False writes no beams: do not pass out")
            ld = _padded_pitch(self.out_shape[2])
            probe = _lib.PwdShape.from_buffer_copy(self.shape)
            probe.out_ld = ld
            if not self.lib.sfa_pwd_is_fused(ctypes.byref(probe)):
                raise ValueError("a power map without the beams needs the fused steering kernel (complex64, steering "
                                 "form, <= 64 microphones and <= 64 filter columns)")
        else:
            if out is None:
                out = _alloc_out(*self.out_shape, self.cdtype, self.dev)
            ld = _check_out(out, *self.out_shape, self.cdtype, self.dev)
            self.shape.out_pad_writable = 1 if (_pad_owned(out) if pad_writable is None else pad_writable) else 0
        # the plan (kernel choice, workspace layout) depends on the row pitch of `out`: prepare again if it changed
        if self._prepared_ld is None or ld != self.shape.out_ld:
            self.shape.out_ld = ld
            if self.lib.sfa_pwd_work_bytes(ctypes.byref(self.shape)) > self.nbytes:
                self.nbytes = self.lib.sfa_pwd_work_bytes(ctypes.byref(self.shape))
                self.work = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
            self.prepare()
        pm = None
        if power is not None and power is not False:
            if self.precision == "c128":
                raise ValueError("the fused power map is available for the complex64 precisions")
            F, L = self.out_shape[0], self.out_shape[1]
            if power is True:
                if self.power is None:
                    self.power = torch.empty((F, L), dtype=torch.float32, device=self.dev)
                pm = self.power
            else:
                pm = power
                if tuple(pm.shape) != (F, L) or pm.dtype != torch.float32 or not pm.is_contiguous() or pm.device != self.dev:
                    raise ValueError("power must be a contiguous float32 tensor of shape (F, L)")
                self.power = pm
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.sfa_pwd_run(ctypes.byref(self.shape), _lib.ptr(self._Y_look), _lib.ptr(self.dn),
                                            _lib.ptr(X), _lib.ptr(out), _lib.ptr(pm), _lib.ptr(self.work),
                                            self.nbytes, _lib.stream_ptr(self.dev)))
        return out if beams else pm


class PwdHost:
    """Host-buffer plane-wave decomposition through the C ABI (``sfa_pwd_host_*``):
    NumPy arrays in, NumPy array out; the H2D copy of X, the kernels and the D2H copy of the beams
    are pipelined over frequency chunks inside the library."""

    def __init__(self, Y_look, dn, Y_mic, T, precision="c64", form="auto", device=0, chunk_bins=0):
        self.lib = _lib.require_cuda()
        self.np_dtype = np.complex128 if precision == "c128" else np.complex64
        Y_look = np.ascontiguousarray(Y_look, dtype=np.complex128)
        Y_mic = np.ascontiguousarray(Y_mic, dtype=np.complex128)
        dn = np.ascontiguousarray(dn, dtype=np.complex128)
        if dn.ndim == 1:
            dn = dn[None, :]
        F, Q = dn.shape[0], Y_mic.shape[0]

        class _S:                                   # duck-typed for _shape_of
            pass
        self.shape = _shape_of(torch.empty(Y_look.shape, device="meta"), torch.empty(dn.shape, device="meta"),
                               torch.empty(Y_mic.shape, device="meta"), F, Q, int(T), precision, form)
        self.ctx = ctypes.c_void_p()
        _lib.check(self.lib.sfa_pwd_host_create(ctypes.byref(self.ctx), ctypes.byref(self.shape), int(device),
                                                int(chunk_bins)))
        self._pinned = []
        try:
            _lib.check(self.lib.sfa_pwd_host_set_operators(self.ctx, Y_look.ctypes.data, Y_mic.ctypes.data,
                                                           dn.ctypes.data))
        except Exception:
            self.close()                         # e.g. SFA_ERR_RANGE: do not leak the context and its device buffers
            raise
        self.in_shape = (F, Q, int(T))
        self.out_shape = (F, Y_look.shape[0], int(T))

    def set_operators(self, Y_look, dn, Y_mic):
        """Replace the resident operators.  CUDA tensors (the outputs of `angular.sht_matrix` /
        `radial.radial_filters`) are copied device to device; NumPy arrays go through the host path."""
        if all(isinstance(a, torch.Tensor) and a.is_cuda for a in (Y_look, dn, Y_mic)):
            Y_look, dn, Y_mic = (as_complex_tensor(a, a.device) for a in (Y_look, dn, Y_mic))
            _check_c64_range(dn, "c128" if self.np_dtype == np.complex128 else "c64")
            dev = Y_look.device
            with torch.cuda.device(dev):
                _lib.check(self.lib.sfa_pwd_host_set_operators_dev(self.ctx, _lib.ptr(Y_look), _lib.ptr(Y_mic),
                                                                   _lib.ptr(dn), _lib.stream_ptr(dev)))
            return
        Y_look = np.ascontiguousarray(Y_look, dtype=np.complex128)
        Y_mic = np.ascontiguousarray(Y_mic, dtype=np.complex128)
        dn = np.ascontiguousarray(dn, dtype=np.complex128)
        _lib.check(self.lib.sfa_pwd_host_set_operators(self.ctx, Y_look.ctypes.data, Y_mic.ctypes.data, dn.ctypes.data))

    def pinned_empty(self, shape, dtype=None):
        """Page-locked NumPy array (cudaHostAlloc) for X / out: makes the copies asynchronous."""
        dt = np.dtype(self.np_dtype if dtype is None else dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = ctypes.c_void_p()
        _lib.check(self.lib.sfa_host_alloc(ctypes.byref(p), max(n, 1)))
        self._pinned.append(p)
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def power_map(self, X, power=None, out=None):
        """Beam power map mean_t |q|^2 ([F, L] float32) through the host path; the beams themselves are only
        copied back when `out` is given."""
        X = np.ascontiguousarray(X, dtype=self.np_dtype)
        if X.ndim == 2:
            X = X[:, :, None]
        if X.shape != self.in_shape:
            raise ValueError("X must have shape {}".format(self.in_shape))
        if power is None:
            power = np.empty(self.out_shape[:2], dtype=np.float32)
        elif power.shape != self.out_shape[:2] or power.dtype != np.float32 or not power.flags.c_contiguous:
            raise ValueError("power must be a C-contiguous float32 array of shape {}".format(self.out_shape[:2]))
        if out is not None and (out.shape != self.out_shape or out.dtype != self.np_dtype or not out.flags.c_contiguous):
            raise ValueError("out must be a C-contiguous {} array of shape {}".format(self.np_dtype, self.out_shape))
        _lib.check(self.lib.sfa_pwd_host_run_power(self.ctx, X.ctypes.data, out.ctypes.data if out is not None else None,
                                                   power.ctypes.data))
        return power

    def __call__(self, X, out=None):
        X = np.ascontiguousarray(X, dtype=self.np_dtype)
        if X.ndim == 2:
            X = X[:, :, None]
        if X.shape != self.in_shape:
            raise ValueError("X must have shape {}".format(self.in_shape))
        if out is None:
            out = np.empty(self.out_shape, dtype=self.np_dtype)
        elif out.shape != self.out_shape or out.dtype != self.np_dtype or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous {} array of shape {}".format(self.np_dtype, self.out_shape))
        _lib.check(self.lib.sfa_pwd_host_run(self.ctx, X.ctypes.data, out.ctypes.data))
        return out

    def close(self):
        if self.ctx:
            self.lib.sfa_pwd_host_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()
        for p in self._pinned:
            self.lib.sfa_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
