"""sfa_numpy_b200 -- B200 (sm_100a) implementation of the sfa-numpy `micarray.modal` hot path.

Drop-in use:  ``import sfa_numpy_b200 as micarray`` then
``micarray.modal.angular.sht_matrix(...)``, ``micarray.modal.radial.spherical_pw(...)`` ...
Functions take NumPy arrays / scalars / torch tensors and return torch CUDA tensors
(complex128 / float64, same shapes, squeeze rules and exceptions as the reference).
All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/sfa_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import modal
from . import util
from . import fft
from . import pipeline
