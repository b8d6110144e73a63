"""CPU: the C ABI library builds, loads, and exports every symbol include/sfa_b200.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from sfa_numpy_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sfa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfa_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from sfa_numpy_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libsfa_b200.so does not export " + name
        assert name in _lib.SIGNATURES, "no ctypes signature for " + name
    assert set(_lib.SIGNATURES) == set(names)


def test_status_strings_and_planning(lib):
    from sfa_numpy_b200 import _lib
    assert lib.sfa_version() >= 100
    assert lib.sfa_strerror(0) == b"ok"
    assert b"workspace" in lib.sfa_strerror(-4)
    assert b"float32 range" in lib.sfa_strerror(-6)
    s = _lib.PwdShape(1024, 2562, 64, 938, 81, 9, 1, 0, 0)
    assert 0 < lib.sfa_pwd_work_bytes(ctypes.byref(s)) < 1 << 30
    bad = _lib.PwdShape(4, 8, 8, 8, 80, 9, 1, 0, 0)           # 9*9 != 80
    assert lib.sfa_pwd_work_bytes(ctypes.byref(bad)) == 0
    assert lib.sfa_radial_work_bytes(1, 8, 1024) > 0


def test_sass_has_tcgen05_and_tma():
    """The stage-3 kernel really is a Blackwell tensor-core kernel: UTC*MMA (tcgen05.mma),
    LDTM/STTM (tcgen05.ld/st) and UTMALDG (TMA) appear in the SASS of the built library."""
    import shutil
    import subprocess
    from sfa_numpy_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass or (mnemonic == "UTCHMMA" and "UTC" in sass and "MMA" in sass), mnemonic


def test_no_cpu_fallback_without_gpu():
    import torch
    import sfa_numpy_b200 as micarray
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        micarray.modal.angular.sht_matrix(2, [0.1], [0.2])
    with pytest.raises(RuntimeError):
        micarray.modal.radial.weights(2, [0.1], "open")


def test_product_never_imports_oracle():
    """The package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "sfa_numpy_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("oracle/", "oracle/") or f == "build.py" or \
                    "import oracle" not in src and "from oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src, f


def test_pwd_workspace_plan_without_gpu():
    """sfa_pwd_work_bytes runs the planner (form selection, chunking, scratch of the K-loop kernel) on the
    host: every BASELINE shape gets a finite, plausible workspace in each form it supports."""
    from sfa_numpy_b200 import _lib
    lib = _lib.load()
    shapes = {  # F, L, Q, T, M, Cd, per_order
        "cfg1": (512, 360, 100, 256, 25, 5, 1),
        "cfg2": (1024, 360, 32, 256, 33, 33, 0),
        "cfg3": (1024, 2562, 64, 938, 81, 9, 1),
        "cfg4": (1024, 2562, 578, 64, 289, 17, 1),
        "cfg5": (4096, 16384, 1089, 32, 1089, 33, 1),
    }
    for name, (F, L, Q, T, M, Cd, per) in shapes.items():
        for prec in (0, 1, 2):
            sizes = {}
            for form in (0, 1, 2):
                sh = _lib.PwdShape(F, L, Q, T, M, Cd, per, prec, form, 0)
                sizes[form] = lib.sfa_pwd_work_bytes(sh)
            assert 0 < sizes[0] < 8 << 30, (name, prec, sizes)          # auto: always plannable, a few GB at most
            assert sizes[2] > 0, (name, prec)                            # factored form: always
            assert sizes[0] in (sizes[1], sizes[2]), (name, prec, sizes)  # auto picks one of the two forms
    # malformed shapes are rejected (0 bytes), not planned
    assert lib.sfa_pwd_work_bytes(_lib.PwdShape(8, 4, 4, 4, 9, 4, 1, 0, 0, 0)) == 0      # Cd^2 != M
    assert lib.sfa_pwd_work_bytes(_lib.PwdShape(8, 4, 4, 4, 9, 3, 1, 7, 0, 0)) == 0      # unknown precision


def test_grid_lebedev_surface():
    """angular.grid_lebedev exists with the reference's error behaviour (quadpy is optional there too)."""
    import pytest
    from sfa_numpy_b200.modal import angular
    with pytest.raises(ValueError):
        angular.grid_lebedev(66)
    try:
        import quadpy  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            angular.grid_lebedev(4)


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/sfa_b200.h must compile as C (the drop-in boundary is a C ABI) and the ctypes mirror of
    sfa_pwd_shape must have the compiler's layout: a field added on one side only would shift every argument."""
    import shutil
    import subprocess
    from sfa_numpy_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    fields = [name for name, _ in _lib.PwdShape._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "sfa_b200.h"', 'int main(void) {',
            '  printf("size %zu\\n", sizeof(sfa_pwd_shape));']
    prog += ['  printf("%s %%zu\\n", offsetof(sfa_pwd_shape, %s));' % (f, f) for f in fields]
    prog += ['  printf("c128 %zu c64 %zu\\n", sizeof(sfa_c128), sizeof(sfa_c64));', '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "size %d" % ctypes.sizeof(_lib.PwdShape)
    for line, f in zip(out[1:], fields):
        assert line == "%s %d" % (f, getattr(_lib.PwdShape, f).offset), line
    assert out[1 + len(fields)] == "c128 16 c64 8"
