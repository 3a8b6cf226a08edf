"""Builds devicekmc_b200/lib/libdkmc_b200.so — the C-ABI library (include/dkmc.h) with every
hand-written sm_100a kernel.  In-tree, explicit nvcc, no JIT cache: the .so travels with the
repo snapshot to the GPU box.  `python -m devicekmc_b200.build [--force]`.
"""
from __future__ import annotations

import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIBDIR = os.path.join(_PKG, "lib")
LIB = os.path.join(LIBDIR, "libdkmc_b200.so")
INCLUDE = os.path.join(_ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-extended-lambda",
          "-I", INCLUDE, "-I", CSRC] + os.environ.get("DKMC_EXTRA_NVCC", "").split()   # experiments only

# per-file extra flags: events.cu keeps the reference's x86-64 rounding (no FMA contraction)
SOURCES = {
    "context.cu": [],
    "graph.cu": [],
    "solver.cu": [],
    "pairwise.cu": [],
    "events.cu": ["-fmad=false"],
    "probes.cu": [],
}


def _newer(a: str, b: str) -> bool:
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "dkmc.h"))
    objs = []
    relink = force or not os.path.exists(LIB)
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            relink = True
    if relink:
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lnccl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
