"""In-tree build of libb2piv.so (nvcc, sm_100a only).  Cross-compiles without a GPU."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "b2piv.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))] + [
    os.path.join(HERE, "..", "include", "b2piv.h")
]
LIB = os.path.join(HERE, "libb2piv.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libb2piv.so cannot be built (pyorc_b200 has no CPU fallback)")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile pyorc_b200/libb2piv.so if missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, SRC, "-o", LIB, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
