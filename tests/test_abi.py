"""CPU: the C-ABI library builds, loads, and exports every symbol include/b2piv.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from pyorc_b200 import engine

    return engine.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2piv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2piv_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from pyorc_b200 import engine

    assert header_symbols() == sorted(engine.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.b2piv_version() >= 100


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.b2piv_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.b2piv_last_error(None)
    from pyorc_b200.engine import Engine

    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyorc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/ffpiv_oracle.py)", ""), f


def _build_c_example(tmp_path):
    import subprocess

    exe = str(tmp_path / "piv_pairs")
    libdir = os.path.join(ROOT, "pyorc_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "piv_pairs.c"), "-o", exe, "-L", libdir, "-l:libb2piv.so", "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_example_builds_links_and_fails_loudly_without_a_gpu(lib, tmp_path):
    """include/b2piv.h is a C header (C99, -pedantic -Werror) and the library links into a plain C program - no Python, no torch
    types at the boundary (examples/piv_pairs.c).  Without a GPU the program reports the ABI's error and status code; with one it
    must return what the ctypes binding returns, bit for bit."""
    import subprocess

    import numpy as np
    import torch

    exe = _build_c_example(tmp_path)
    usage = subprocess.run([exe], capture_output=True, text=True)
    assert usage.returncode == 64 and "ABI version" in usage.stderr
    from pyorc_b200 import synth

    imgs = synth.particle_frames(3, 100, 140, dtype=np.uint8)
    raw, out = str(tmp_path / "frames.raw"), str(tmp_path / "out.raw")
    imgs.tofile(raw)
    res = subprocess.run([exe, raw, "3", "100", "140", "32", "32", "16", "16", out], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert res.returncode == 2                                   # B2PIV_ERR_CUDA
        assert "no CPU fallback" in res.stderr and not os.path.exists(out)
        return
    assert res.returncode == 0, res.stderr
    from pyorc_b200.engine import Engine

    with Engine(0) as e:
        want = np.stack(e.pairs(imgs, (32, 32), (16, 16)))
    got = np.fromfile(out, np.float32).reshape(want.shape)
    assert np.array_equal(got, want, equal_nan=True)
