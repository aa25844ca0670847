"""Test configuration.

* ``-m "not gpu"`` (CPU container): oracle vs golden vectors, host logic, C-ABI
  symbol checks.  ``-m gpu`` (B200 box): parity of the CUDA path against the
  oracle, always through the C ABI (ctypes).
* The oracle (``oracle/``) and the test-only ``ase`` shim are put on ``sys.path``
  here and nowhere else; the shim only when a real ``ase`` is not importable.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if importlib.util.find_spec('ase') is None:
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ase_shim'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def lj_engine():
    from mcpy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope='session')
def emt_engine():
    from mcpy_b200 import Engine
    e = Engine(0)
    e.set_emt()
    yield e
    e.close()


@pytest.fixture(scope='session')
def fv_engine():
    from mcpy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()
