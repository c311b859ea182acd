"""Build the in-tree CUDA library (bipymc_b200/lib/libbipymc_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The library links cudart statically, so it loads on
a CPU-only box too (the CPU test-suite checks its exported symbols).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libbipymc_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["rng.cuh", "step.cuh", "targets.cuh", "kernels_generic.cuh", "kernels_fused.cuh", "kernels_fused_v4.cuh", "diagnostics.cuh", "gauss_dmma.cuh", "sync.cuh",
           os.path.join("..", "..", "include", "bipymc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.basename(__file__)]:
        p = os.path.join(CSRC, f) if f.endswith((".cu", ".cuh", ".h")) else os.path.join(HERE, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if not force and os.path.exists(LIB):
            return LIB          # a prebuilt library travelled with the tree; nothing to rebuild it with
        raise RuntimeError("nvcc not found; cannot build %s" % LIB)
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
