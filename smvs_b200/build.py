"""Builds smvs_b200/libsmvs_b200.so (sm_100a only) with nvcc, in-tree.

nvcc cross-compiles without a GPU. The library links cudart statically and
has no other dependency, so the built file travels as-is.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmvs_b200.so")
SOURCES = ["api.cu", "gn_construct.cu", "cg.cu", "update.cu", "sgm.cu", "views.cu",
           "visibility.cu", "microbench.cu", "topology.cu", "cut_maps.cu"]
HEADERS = ["common.cuh", "gn_math.cuh", "patch_eval.cuh", os.path.join("..", "..", "include", "smvs_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
    "--shared", "-cudart", "static",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-ccbin", "g++"] + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsmvs_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
