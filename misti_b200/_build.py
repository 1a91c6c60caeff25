"""Build libmisti_b200.so in-tree with nvcc for sm_100a (no CPU variant exists).

Used by ``__graft_entry__.build()`` and by ``python -m misti_b200._build``.  The built library is
git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmisti_b200.so")
SOURCES = [os.path.join(CSRC, "misti_kernels.cu")]
HEADERS = [os.path.join(CSRC, n) for n in ("misti_math.cuh", "misti_model.cuh", "misti_jsfs.cuh", "misti_optim.cuh", "misti_tables.h", "misti_pair.cuh", "misti_pair_code.h")] + [
    os.path.join(ROOT, "include", "misti_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmisti_b200.so cannot be built (there is no CPU build of this library)")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    if not force and not stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
