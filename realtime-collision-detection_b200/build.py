"""Build librcd_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

    python realtime-collision-detection_b200/build.py [--force] [--verbose]

The shared library is self-contained (static cudart, no torch dependency) and travels to the
GPU box with the snapshot.  There is no other build product and no fallback path.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librcd_b200.so")
SOURCES = ["rcd_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]  # every source and header under csrc/
    deps += [os.path.join(os.path.dirname(HERE), "include", "rcd.h"), os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    tmp = LIB + ".building"  # (renamed into place when complete: a snapshot never sees half a library)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    # the host compiler must be one nvcc 12.9 accepts; /usr/bin/g++ (13.x) is
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building librcd_b200.so")
    os.replace(tmp, LIB)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
