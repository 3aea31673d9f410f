"""In-tree build of the CUDA library (sm_100a only) with nvcc.  No JIT cache, no torch extension:
the product is a plain C-ABI shared object next to this file so that it travels with the repo."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libernet_b200.so")
ROOT_INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v",
]


def sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    out.append(os.path.join(ROOT_INCLUDE, "ernet_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into libernet_b200.so.  Raises if nvcc is missing or compilation fails."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libernet_b200.so (there is no CPU fallback)")
    cus = [s for s in sources() if s.endswith(".cu")]
    extra = os.environ.get("ERNET_NVCC_EXTRA", "").split()          # study builds, e.g. -DERNET_TIMELINE
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", ROOT_INCLUDE, "-o", LIB + ".tmp"] + cus
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:])
        print(r.stderr[-8000:])
    if r.returncode:
        raise RuntimeError("nvcc failed building libernet_b200.so")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
