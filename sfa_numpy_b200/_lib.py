"""ctypes binding of libsfa_b200.so (the C ABI declared in include/sfa_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is
present, every compute entry point raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsfa_b200.so")
# tuning experiments: an alternative build of the same library (tools/ only)
LIB_PATH = os.environ.get("SFA_B200_LIB", LIB_PATH)

c_void_p, c_int, c_int64, c_double, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                               ctypes.c_double, ctypes.c_size_t)


class PwdShape(ctypes.Structure):
    """sfa_pwd_shape of include/sfa_b200.h."""
    _fields_ = [("F", c_int64), ("L", c_int64), ("Q", c_int64), ("T", c_int64),
                ("M", c_int), ("Cd", c_int), ("d_is_per_order", c_int),
                ("precision", c_int), ("form", c_int), ("out_ld", c_int64), ("out_pad_writable", c_int)]


# name -> (restype, argtypes); must list every symbol of include/sfa_b200.h
SIGNATURES = {
    "sfa_strerror": (ctypes.c_char_p, [c_int]),
    "sfa_last_cuda_error": (ctypes.c_char_p, []),
    "sfa_version": (c_int, []),
    "sfa_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "sfa_set_option": (c_int, [ctypes.c_char_p, c_double]),
    "sfa_get_option": (c_double, [ctypes.c_char_p]),
    "sfa_launch_count": (ctypes.c_longlong, []),
    "sfa_prof_enable": (c_int, [c_int]),
    "sfa_prof_read": (c_int, [ctypes.POINTER(c_double), ctypes.POINTER(ctypes.c_longlong)]),
    "sfa_prof_span": (c_int, [ctypes.POINTER(c_double)]),
    "sfa_prof_timeline": (c_int, [ctypes.POINTER(c_double), c_int]),
    "sfa_sht_matrix": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "sfa_cht_matrix": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_radial_work_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "sfa_radial_filter": (c_int, [c_int, c_int, c_int64, c_void_p, c_double, c_double, c_int, c_int, c_int,
                                  c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_replace_zeros": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sfa_regularize": (c_int, [c_int64, c_void_p, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "sfa_repeat_n_m": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "sfa_diagonal_stack": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "sfa_legendre_matrix": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "sfa_grid_points": (c_int64, [c_int, c_int]),
    "sfa_grid": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_peer_copy": (c_int, [c_void_p, c_int, c_void_p, c_int, c_size_t, c_void_p]),
    "sfa_matdiagmul": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_norm_of_columns": (c_int, [c_int64, c_int64, c_void_p, c_double, c_void_p, c_void_p]),
    "sfa_coherence_of_columns": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_pwd_work_bytes": (c_size_t, [ctypes.POINTER(PwdShape)]),
    "sfa_pwd_apply": (c_int, [ctypes.POINTER(PwdShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "sfa_pwd_prepare": (c_int, [ctypes.POINTER(PwdShape), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sfa_pwd_is_fused": (c_int, [ctypes.POINTER(PwdShape)]),
    "sfa_pwd_run": (c_int, [ctypes.POINTER(PwdShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_size_t, c_void_p]),
    "sfa_cgemm3x": (c_int, [c_int64] * 4 + [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                            c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "sfa_pwd_host_create": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(PwdShape), c_int, c_int64]),
    "sfa_pwd_host_set_operators": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_pwd_host_run": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sfa_pwd_host_set_operators_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_pwd_host_run_power": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfa_pwd_host_destroy": (c_int, [c_void_p]),
    "sfa_power_map": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "sfa_irfft_work_bytes": (c_size_t, [c_int64, c_int]),
    "sfa_irfft_axis0": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                c_void_p, c_size_t, c_void_p]),
    "sfa_stft_work_bytes": (c_size_t, [c_int, c_int]),
    "sfa_stft": (c_int, [c_int64, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                         c_void_p, c_size_t, c_void_p]),
    "sfa_host_alloc": (c_int, [ctypes.POINTER(c_void_p), c_size_t]),
    "sfa_host_free": (c_int, [c_void_p]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libsfa_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class SfaError(RuntimeError):
    pass


def check(status):
    if status != 0:
        lib = load()
        msg = lib.sfa_strerror(status).decode()
        if status == -2:
            msg += ": " + lib.sfa_last_cuda_error().decode()
        raise SfaError("libsfa_b200: %s (status %d)" % (msg, status))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("sfa_numpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load()


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
