"""In-tree build of libb2piv.so (nvcc, sm_100a only).  Cross-compiles without a GPU.

One object per translation unit of ``csrc/`` (``k_*.cu``: one kernel family each; ``abi_*.cu``: the C ABI), compiled in
parallel into ``pyorc_b200/build/`` and linked into ``pyorc_b200/libb2piv.so``.  A unit is rebuilt when it, any header of
``csrc/`` or ``include/b2piv.h`` is newer than its object.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb2piv.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return sorted(hs) + [os.path.join(HERE, "..", "include", "b2piv.h")]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libb2piv.so cannot be built (pyorc_b200 has no CPU fallback)")
    return nvcc


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale_units():
    ht = max(os.path.getmtime(h) for h in headers() if os.path.exists(h))
    out = []
    for s in sources():
        o = _obj(s)
        if not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), ht):
            out.append(s)
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + headers()
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    """Compile pyorc_b200/libb2piv.so if missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    units = sources() if force else _stale_units()

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", _obj(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    jobs = jobs or min(len(units), os.cpu_count() or 1) or 1
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for src, res in ex.map(compile_one, units):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout + res.stderr)
            if verbose:
                print(f"== {os.path.basename(src)}\n{res.stderr}")
    objs = [_obj(s) for s in sources()]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", *objs, "-o", LIB,
           "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv or len(sys.argv) == 1, verbose=True))
