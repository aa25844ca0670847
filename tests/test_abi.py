"""CPU: the C-ABI library loads and exports exactly what include/mcpy_b200.h declares.
(No compute calls: there is no GPU in the build container.)"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'mcpy_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mcpyb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from mcpy_b200 import _lib, build
    build.build()                       # cross-compiles for sm_100a without a GPU
    lib = _lib.load()
    declared = _header_symbols()
    assert declared, 'no symbols parsed from the header'
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared


def test_version_string():
    from mcpy_b200 import _lib
    assert b'sm_100a' in _lib.load().mcpyb_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from mcpy_b200 import Engine, EngineError
    from mcpy_b200.calculators import B200Calculator
    with pytest.raises(EngineError):
        Engine(0)
    with pytest.raises(EngineError):
        B200Calculator('lj')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'mcpy_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_pinned_allocation_is_declared_and_exported():
    # mcpyb_alloc_pinned needs a CUDA runtime context to succeed; here only the binding is checked
    from mcpy_b200 import _lib
    lib = _lib.load()
    assert lib.mcpyb_alloc_pinned.restype is not None and hasattr(lib, 'mcpyb_free_pinned')
