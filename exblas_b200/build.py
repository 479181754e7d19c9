"""Build the in-tree CUDA library (sm_100a only).  Used by __graft_entry__.build() and by hand:

    python -m exblas_b200.build [--force]

The .so is written next to this file (exblas_b200/libexblas_b200.so) so that it travels to the
GPU box with the repository snapshot; it is git-ignored.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libexblas_b200.so")
SOURCES = [os.path.join(CSRC, "exblas_b200.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("reduce_kernel.cuh", "superacc.cuh", "window.cuh", "gemv_kernel.cuh",
                                                   "segments_kernel.cuh", "microbench.cuh")] + [os.path.join(HERE, "..", "include", "exblas_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=true",          # TwoSum / TwoProd use __dadd_rn/__dmul_rn/__fma_rn, which are never contracted
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",    # share the CUDA runtime (and its primary context) with torch in the same process
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-ldl"]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that lack OpenMP specs; nvcc only needs a plain host g++
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("nvcc failed building libexblas_b200.so")
    if verbose:
        sys.stderr.write(proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
