"""Build libsfa_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the built
library travels with the repository snapshot to the GPU box."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libsfa_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


STAMP = os.path.join(HERE, "lib", "build.stamp")


def source_hash():
    """Content hash of everything the library is built from (robust to mtime changes when the tree is
    copied to another machine)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(b"debug" if os.environ.get("SFA_BUILD_DEBUG_HOOKS") else b"")
    h.update(os.environ.get("SFA_BUILD_DEFS", "").encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_variant(name, defs, verbose=False):
    """Extra build of the library with additional -D flags (tools/ experiments only; never loaded by the
    package unless SFA_B200_LIB points at it)."""
    out = os.path.join(HERE, "lib", "libsfa_b200_%s.so" % name)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + list(defs) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + out)
    return out


def build_debug(verbose=False):
    """Build with the timing-attribution hooks compiled in."""
    return build_variant("dbg", ["-DSFA_DEBUG_HOOKS"], verbose)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    extra = ["-DSFA_DEBUG_HOOKS"] if os.environ.get("SFA_BUILD_DEBUG_HOOKS") else []
    extra += os.environ.get("SFA_BUILD_DEFS", "").split()          # tuning experiments, e.g. -DSFA_GEMM_STAGES=2
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsfa_b200.so")
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
