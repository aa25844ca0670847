"""Brute-force fp64 pair enumeration (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates what ASE 3.23.0 ``ase/neighborlist.py::PrimitiveNeighborList.build``
delivers to the calculators the reference uses (SURVEY.md section 8a rows
a8/a10): the set of ``(i, j, S)`` -- atom, neighbour, integer cell offset --
such that the vector the calculators form,

    d = positions[j] + S @ cell - positions[i]

(``ase/calculators/lj.py``: ``positions[neighbors] + cells - positions[ii]``;
``ase/calculators/emt.py``: ``positions[a2] + offset - positions[a1]``), is
within the cutoff.  ASE is not vendored in ``/root/reference`` (it is the
third-party dependency ``ase==3.23.0``, ``uv.lock:9-20``) and the reference has
no test that pins a neighbour list, so parity for this file is "unpinned": the
definition below -- O(N^2 * images) enumeration with ASE's expression order,
``(dx*dx + dy*dy) + dz*dz`` in unfused fp64 -- IS the bit-exactness contract the
CUDA cell list is tested against.
"""
from __future__ import annotations

import numpy as np


def complete_cell(cell, pbc):
    """3x3 cell with non-periodic / zero rows replaced by unit-ish vectors so
    that fractional coordinates exist (mirrors ``ase.geometry.complete_cell``
    in intent; only used for image bookkeeping, never for distances)."""
    cell = np.array(cell, dtype=float).reshape(3, 3)
    out = cell.copy()
    for d in range(3):
        if not np.any(out[d]):
            # choose a vector orthogonal to the other two rows
            others = [out[e] for e in range(3) if e != d and np.any(out[e])]
            if len(others) == 2:
                v = np.cross(others[0], others[1])
            elif len(others) == 1:
                trial = np.eye(3)[np.argmin(np.abs(others[0]))]
                v = np.cross(others[0], trial)
            else:
                v = np.eye(3)[d]
            out[d] = v / np.linalg.norm(v)
    return out


def cell_heights(cell):
    """Perpendicular widths of the parallelepiped along each lattice vector."""
    cell = np.asarray(cell, dtype=float)
    vol = abs(np.linalg.det(cell))
    h = np.empty(3)
    for d in range(3):
        a, b = cell[(d + 1) % 3], cell[(d + 2) % 3]
        h[d] = vol / np.linalg.norm(np.cross(a, b))
    return h


def shift_vector(S, cell):
    """``S @ cell`` evaluated left to right, one product per lattice row.

    For orthorhombic cells (every BASELINE config) two of the three products
    are exactly zero, so the result is the single correctly-rounded product
    that any evaluation order gives."""
    return S[0] * cell[0] + S[1] * cell[1] + S[2] * cell[2]


def wrap_counts(positions, cell, pbc):
    """Integer ``w`` per atom and dimension such that ``positions - w @ cell``
    has fractional coordinates in [0, 1) along periodic dimensions."""
    pbc = np.asarray(pbc, dtype=bool).reshape(3)
    if len(positions) == 0:
        return np.zeros((0, 3), dtype=np.int64)
    cc = complete_cell(cell, pbc)
    frac = np.linalg.solve(cc.T, np.asarray(positions, dtype=float).T).T
    w = np.floor(frac).astype(np.int64)
    w[:, ~pbc] = 0
    return w


def neighbor_pairs(positions, cell, pbc, cutoff, *, inclusive=False,
                   use_sqrt=False, half=False, chunk=256):
    """Enumerate every neighbour image within ``cutoff``.

    Parameters
    ----------
    inclusive : ``r2 <= cutoff**2`` (ASE LennardJones keeps pairs unless
        ``r2 > rc**2``) instead of strict ``<``.
    use_sqrt : compare ``sqrt(r2) < cutoff`` (ASE EMT: ``r = sqrt(dot(d, d));
        if r < self.rc_list``) instead of comparing squares.
    half : keep each unordered pair once, as ASE's default half list does
        (``i < j``, or ``i == j`` with the lexicographically positive S).

    Returns ``(i, j, S, r2)`` sorted by ``(i, j, Sx, Sy, Sz)``; ``S`` is int64
    ``(npairs, 3)`` and ``r2`` the fp64 squared distance of that image.
    """
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    cell = np.array(cell, dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(pbc, dtype=bool).reshape(3)
    n = len(pos)
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros((0, 3), dtype=np.int64), np.zeros(0)
    w = wrap_counts(pos, cell, pbc)
    nimg = np.zeros(3, dtype=int)
    if pbc.any():
        h = cell_heights(complete_cell(cell, pbc))
        for d in range(3):
            if pbc[d]:
                nimg[d] = int(np.ceil(cutoff / h[d]))
    rc2 = cutoff * cutoff
    out_i, out_j, out_S, out_r2 = [], [], [], []
    for sx in range(-nimg[0], nimg[0] + 1):
        for sy in range(-nimg[1], nimg[1] + 1):
            for sz in range(-nimg[2], nimg[2] + 1):
                s = np.array([sx, sy, sz], dtype=np.int64)
                for start in range(0, n, chunk):
                    stop = min(n, start + chunk)
                    # S relative to the ORIGINAL (unwrapped) positions
                    S = s[None, None, :] + w[start:stop, None, :] - w[None, :, :]
                    sv = (S[..., 0, None] * cell[0] + S[..., 1, None] * cell[1]
                          + S[..., 2, None] * cell[2])
                    d = (pos[None, :, :] + sv) - pos[start:stop, None, :]
                    r2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) \
                        + d[..., 2] * d[..., 2]
                    if use_sqrt:
                        m = np.sqrt(r2) < cutoff
                    elif inclusive:
                        m = r2 <= rc2
                    else:
                        m = r2 < rc2
                    ii, jj = np.nonzero(m)
                    gi = ii + start
                    keep = ~((gi == jj) & (S[ii, jj] == 0).all(axis=1))
                    ii, jj, gi = ii[keep], jj[keep], gi[keep]
                    out_i.append(gi)
                    out_j.append(jj)
                    out_S.append(S[ii, jj])
                    out_r2.append(r2[ii, jj])
    i = np.concatenate(out_i).astype(np.int64)
    j = np.concatenate(out_j).astype(np.int64)
    S = np.concatenate(out_S).astype(np.int64).reshape(-1, 3)
    r2 = np.concatenate(out_r2)
    if half:
        lexpos = (S[:, 0] > 0) | ((S[:, 0] == 0) & ((S[:, 1] > 0) | (
            (S[:, 1] == 0) & (S[:, 2] > 0))))
        keep = (i < j) | ((i == j) & lexpos)
        i, j, S, r2 = i[keep], j[keep], S[keep], r2[keep]
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    return i[order], j[order], S[order], r2[order]
