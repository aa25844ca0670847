"""Test configuration.

* ``-m "not gpu"`` (CPU container): oracle vs golden vectors, host logic, C-ABI
  symbol checks.  ``-m gpu`` (B200 box): parity of the CUDA path against the
  oracle, always through the C ABI (ctypes).
* The oracle (``oracle/``), the unmodified reference (``baseline/_ref``, installed by
  ``__graft_entry__.build()``; ``/root/reference`` as a fall-back in the build container) and,
  only when a real ``ase`` is not importable, the ``ase`` stand-in (``baseline/ase_shim``) are
  put on ``sys.path`` here (``baseline/refpath.py``).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline import refpath  # noqa: E402

ASE_KIND, MCPY_ROOT = refpath.enable()


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
