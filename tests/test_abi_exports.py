"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/jwas_b200.h declares; without a GPU the compute entry points fail loudly (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jwas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jwas_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import jwas_b200
    from jwas_b200 import _lib
    if not os.path.exists(jwas_b200.SO_PATH):
        from importlib import import_module
        import_module("jwas_b200.build").build()
    L = ctypes.CDLL(jwas_b200.SO_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/jwas_b200.h but not exported"
    assert _lib.lib() is not None


def test_io_library_exports_every_declared_symbol():
    """include/jwas_io.h <-> libjwasio.so (host C, the genotype-file side of the path)."""
    from jwas_b200 import _io
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jwas_io.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(jw(?:io|ann)_[a-z0-9_]+)\s*\(", src)))
    assert len(syms) >= 10 and "jwann_probit_step" in syms
    L = ctypes.CDLL(_io.SO_PATH)
    for s_ in syms:
        assert hasattr(L, s_), f"{s_} declared in include/jwas_io.h but not exported"
    assert _io.lib() is not None


def test_no_cpu_fallback():
    import jwas_b200
    if jwas_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(jwas_b200.JwasError, match="no CPU fallback"):
        jwas_b200.GpuSweeper(np.zeros((4, 2), np.uint8), 8, 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing shipped may include, import, link or load it."""
    pkg = os.path.join(ROOT, "jwas.jl_b200")
    bad = re.compile(r'(#\s*include\s*[<"][^>"]*oracle|^\s*(from|import)\s+\S*oracle|libjwas_oracle|CDLL\([^)]*oracle)', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f
