"""Build recipe: every CUDA source is compiled for sm_100a with nvcc into in-tree shared objects.

    python -m unipeak_b200.build            # library (+ host CLIs when present)

-fmad=false: the reference's sums round every product separately (no FMA); the kernels use
explicit __dmul_rn/__dadd_rn where it matters and this flag covers the rest.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libunipeak_b200.so")
BIN = os.path.join(ROOT, "bin")

NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "unipeak_b200.h")]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [_nvcc(), "-shared"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


def build_host(force: bool = False) -> None:
    """C++ command-line front ends (drop-in bin/regions, bin/strand_shift, bin/tags_in_regions)."""
    if not os.path.isdir(HOST):
        return
    mains = [f for f in sorted(os.listdir(HOST)) if f.startswith("main_") and f.endswith(".cpp")]
    if not mains:
        return
    os.makedirs(BIN, exist_ok=True)
    common = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp") and not f.startswith("main_")]
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".h"))]
    gxx = shutil.which("g++") or "g++"
    for m in mains:
        exe = os.path.join(BIN, m[len("main_"):-len(".cpp")])
        srcs = [os.path.join(HOST, m)] + common
        if not force and _newer(exe, srcs + hdrs + [LIB]):
            continue
        cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe] + srcs + \
              ["-L", HERE, "-lunipeak_b200", "-Wl,-rpath,$ORIGIN/../unipeak_b200", "-lz", "-lpthread"]
        subprocess.check_call(cmd, cwd=ROOT)


def main() -> None:
    force = "--force" in sys.argv
    build_library(force=force, verbose="-v" in sys.argv)
    build_host(force=force)
    print(LIB)


if __name__ == "__main__":
    main()
