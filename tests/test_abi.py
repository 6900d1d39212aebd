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
