"""CPU oracle for the sfa-numpy `micarray.modal` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sfa_numpy_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or as
the CPU baseline -- never as the thing shipped.

Parity pin: the reference ships no tests, golden vectors or known-answer tests
for this path, so the oracle is pinned against outputs of
the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` with the
``sph_harm`` shim) and committed as ``tests/golden/*.npz``.
"""
from . import sfa_oracle  # noqa: F401
