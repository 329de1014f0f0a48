"""Builds librtb200.so (sm_100a only) in-tree with nvcc.  `python build.py [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librtb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC",
          # denormals flushed: __fdividef / rsqrtf become single MUFU instructions
          # (without it each carries a 4-instruction denormal rescue)
          "-ftz=true"]
EXTRA = os.environ.get("RTB_NVCC_EXTRA", "").split()   # experiment knobs, e.g. -DRTB_SHADE_THREADS=160 -DRTB_SHADE_MINB=4
SOURCES = ["engine.cu", "lbvh.cu", "scene_host.cpp", "octree_host.cpp"]


def _deps(src: str) -> list[str]:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "rtb200.h"))
    return [os.path.join(CSRC, src)] + hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    # the objects depend on the flags too: a change of RTB_NVCC_EXTRA alone must rebuild (an experiment that silently measured
    # the previous binary is worse than a slow build)
    stamp, flags = os.path.join(OBJ, "flags.txt"), " ".join([*ARCH, *COMMON, *EXTRA])
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
        with open(stamp, "w") as f:
            f.write(flags)
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        newest = max(os.path.getmtime(d) for d in _deps(src))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            cmd = [NVCC, *ARCH, *COMMON, *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
