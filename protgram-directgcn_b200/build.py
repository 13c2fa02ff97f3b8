"""Builds libpgb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python protgram-directgcn_b200/build.py [--force]

The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpgb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"),
         "-I", os.path.join(HERE, "csrc")]


EXTRA = os.environ.get("PGB200_NVCC_FLAGS", "").split()    # e.g. -DPG_SPMM_MIN_BLOCKS=5 for A/B builds


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + EXTRA + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpgb200.so (see %s)" % log)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
