"""Build recipe for libplenvdb_b200.so (hand-written sm_100a kernels + the C-ABI).

Plain nvcc, one object per source, linked in-tree so the built library travels with the repo snapshot.
Used by ``__graft_entry__.build()`` and runnable as ``python -m plenvdb_b200.build``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libplenvdb_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
          "--expt-relaxed-constexpr"] + os.environ.get("PVDB_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG_DIR, "..", "include", "plenvdb_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
    if not _stale(obj, [src] + _headers()):
        return obj
    cmd = [NVCC] + ARCH_FLAGS + COMMON + ["-Xptxas", "-v"] * bool(verbose) + ["-x", "cu", "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build_library(force=False, verbose=False):
    """Compile every source under csrc/ for sm_100a and link libplenvdb_b200.so. Returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if force or _stale(LIB_PATH, objs):
        cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
