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


def source_hash():
    """sha256 over the names and contents of every source the library is built from."""
    import hashlib
    h = hashlib.sha256()
    for path in sources():
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    return h.hexdigest()


def embedded_hash(lib=LIB):
    """The source hash the library was compiled from (ernet_source_hash() of the C ABI), read from the file itself - no
    dlopen, so it also works where the library cannot be loaded."""
    marker = b"ERNET_SOURCE_HASH="
    try:
        with open(lib, "rb") as f:
            data = f.read()
    except OSError:
        return None
    i = data.find(marker)
    if i < 0:
        return None
    return data[i + len(marker):i + len(marker) + 64].decode("ascii", "replace")


def needs_build():
    """True unless the library on disk was compiled from exactly the sources on disk (hash embedded at compile time;
    file times are not trusted: a shipped prebuilt library must not be reused silently after the sources changed)."""
    return embedded_hash() != source_hash()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into libernet_b200.so.  Raises if nvcc is missing or compilation fails."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libernet_b200.so (there is no CPU fallback)")
    cus = [s for s in sources() if s.endswith(".cu")]
    extra = os.environ.get("ERNET_NVCC_EXTRA", "").split()          # study builds, e.g. -DERNET_TIMELINE
    cmd = [nvcc] + NVCC_FLAGS + extra + [f'-DERNET_SOURCE_HASH="{source_hash()}"', "-I", ROOT_INCLUDE, "-o", LIB + ".tmp"] + cus
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
