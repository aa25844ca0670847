"""Make the unmodified reference (``mcpy``) and an ``ase`` importable.  Used by ``tests/`` and ``bench.py``."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, 'ase_shim')
CANDIDATES = (os.path.join(HERE, '_ref'), '/root/reference')


def enable():
    """Returns ``(ase_kind, mcpy_root)``: ``'ase'`` | ``'shim'`` and the directory ``mcpy`` is imported from
    (None when the reference is nowhere to be found)."""
    kind = 'ase'
    spec = importlib.util.find_spec('ase')
    if spec is None or (spec.origin or '').startswith(SHIM):
        kind = 'shim'
        if SHIM not in sys.path:
            sys.path.insert(0, SHIM)
    root = None
    spec = importlib.util.find_spec('mcpy')
    if spec is not None and spec.origin:
        root = os.path.dirname(os.path.dirname(spec.origin))
    else:
        for cand in CANDIDATES:
            if os.path.isfile(os.path.join(cand, 'mcpy', '__init__.py')):
                sys.path.insert(0, cand)
                root = cand
                break
    return kind, root
