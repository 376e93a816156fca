"""Builds libwignerd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SRC = os.path.join(HERE, "csrc", "wd_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "wd_k2_col.cuh"), os.path.join(HERE, "csrc", "wd_kernels.cuh"), os.path.join(HERE, "csrc", "wd_k2_dmma.cuh"), os.path.join(HERE, "csrc", "wd_k2_stream.cuh"), os.path.join(HERE, "csrc", "wd_k3_tma.cuh"), os.path.join(HERE, "csrc", "wd_rotate.cuh"), os.path.join(ROOT, "include", "wignerd_b200.h")]
LIB = os.path.join(HERE, "libwignerd_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
