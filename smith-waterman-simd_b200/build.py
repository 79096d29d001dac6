"""Build recipe for libswb200.so (the C-ABI shared library with the sm_100a kernels).

    python smith-waterman-simd_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is built IN-TREE next to this file so that it
travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libswb200.so")
SOURCES = [os.path.join(CSRC, "swb200_api.cu"), os.path.join(CSRC, "pairgen.cpp"), os.path.join(CSRC, "hostpack.cpp"), os.path.join(CSRC, "hostprobe.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("sw_core.cuh", "sw_kernel.cuh", "sg_kernel.cuh", "sg2_core.cuh", "sg_host.inc", "sg_pipe.inc", "sg_abi.inc", "sw_params.h", "sw_feed_kernel.cuh", "sw_pair_kernel.cuh", "feed.inc", "pairpath.inc", "peakprobe.inc")] + [
    os.path.join(ROOT, "include", "swb200.h"), os.path.abspath(__file__)]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
