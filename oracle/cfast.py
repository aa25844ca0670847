"""ctypes face of the plain-C oracle ``oracle/c/oracle.c`` (TEST INFRASTRUCTURE).

Same definitions as the numpy oracle, fast enough for the 2 000- and
10 000-atom configurations.  ``tests/test_oracle_lj.py`` / ``test_oracle_emt.py`` pin it to the numpy
restatements, which in turn are pinned to the reference's golden values.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import emt as _emt

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, 'liboracle_c.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'c', 'oracle.c')
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', os.path.join(_HERE, 'c'), '-B'], check=True,
                       capture_output=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_LIB)
        v = C.c_void_p
        lib.oracle_neighbor_pairs.argtypes = [v, C.c_int, v, v, C.c_double, C.c_int, C.c_int64, v, v, v, v]
        lib.oracle_neighbor_pairs.restype = C.c_int64
        lib.oracle_lj_energy.argtypes = [v, C.c_int, v, v, C.c_double, C.c_double, C.c_double, v, v]
        lib.oracle_emt_energy.argtypes = [v, v, C.c_int, v, v, v, C.c_double, C.c_double, C.c_double,
                                          C.c_double, v, v, v]
        lib.oracle_fv_covered.argtypes = [v, C.c_int64, v, v, C.c_int64, v]
        lib.oracle_fv_covered.restype = C.c_int64
        lib.oracle_set_bruteforce.argtypes = [C.c_int]
        _lib = lib
    return _lib


def set_bruteforce(on):
    """Plain O(N^2 * images) loops instead of the bin-pruned ones (same visiting order, same bits)."""
    _load().oracle_set_bruteforce(1 if on else 0)


def _prep(positions, cell, pbc):
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(9)
    pbc = np.ascontiguousarray(np.asarray(pbc, dtype=bool).reshape(3), dtype=np.uint8)
    return pos, cell, pbc


def neighbor_pairs(positions, cell, pbc, cutoff, mode=0):
    """mode 0: r2 < rc^2, 1: r2 <= rc^2, 2: sqrt(r2) < rc.  Sorted like oracle.neighbors."""
    lib = _load()
    pos, cell, pbc = _prep(positions, cell, pbc)
    n = len(pos)
    cap = max(1024, 200 * n)
    while True:
        i = np.empty(cap, np.int32); j = np.empty(cap, np.int32)
        S = np.empty((cap, 3), np.int32); r2 = np.empty(cap)
        k = lib.oracle_neighbor_pairs(pos.ctypes.data, n, cell.ctypes.data, pbc.ctypes.data, float(cutoff),
                                      int(mode), cap, i.ctypes.data, j.ctypes.data, S.ctypes.data, r2.ctypes.data)
        if k < 0:
            raise ValueError('singular cell / periodic axis without lattice vector')
        if k <= cap:
            break
        cap = int(k)
    i, j, S, r2 = i[:k].astype(np.int64), j[:k].astype(np.int64), S[:k].astype(np.int64), r2[:k]
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    return i[order], j[order], S[order], r2[order]


def lj_energy(positions, cell, pbc, sigma=1.0, epsilon=1.0, rc=None, return_pairs=False):
    lib = _load()
    rc = 3 * sigma if rc is None else rc
    pos, cell, pbc = _prep(positions, cell, pbc)
    E = C.c_double(0.0); npairs = C.c_int64(0)
    if lib.oracle_lj_energy(pos.ctypes.data, len(pos), cell.ctypes.data, pbc.ctypes.data, float(sigma),
                            float(epsilon), float(rc), C.byref(E), C.byref(npairs)):
        raise ValueError('singular cell / periodic axis without lattice vector')
    return (E.value, npairs.value) if return_pairs else E.value


def emt_energy(positions, numbers, cell, pbc, return_details=False):
    lib = _load()
    pos, cell, pbc = _prep(positions, cell, pbc)
    Zs, table = _emt.parameter_table()
    slot_of = {int(z): k for k, z in enumerate(Zs)}
    slot = np.array([slot_of[int(z)] for z in np.asarray(numbers).reshape(-1)], dtype=np.int32)
    rc, rc_list, acut = _emt.cutoffs()
    table = np.ascontiguousarray(table)
    sig = np.zeros(max(1, len(pos)))
    E = C.c_double(0.0); npairs = C.c_int64(0)
    if lib.oracle_emt_energy(pos.ctypes.data, slot.ctypes.data, len(pos), cell.ctypes.data, pbc.ctypes.data,
                             table.ctypes.data, rc, rc_list, acut, _emt.BETA, C.byref(E), sig.ctypes.data,
                             C.byref(npairs)):
        raise ValueError('singular cell / periodic axis without lattice vector')
    if return_details:
        return E.value, sig[:len(pos)], npairs.value
    return E.value


def fv_covered(points, positions, radii):
    lib = _load()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    rad = np.ascontiguousarray(radii, dtype=np.float64).reshape(-1)
    cov = np.zeros(len(pts), np.uint8)
    lib.oracle_fv_covered(pts.ctypes.data, len(pts), pos.ctypes.data, rad.ctypes.data, len(pos), cov.ctypes.data)
    return cov.astype(bool)
