"""``find_mic`` / ``wrap_positions`` stand-ins (behaviour of ASE 3.23 geometry)."""
import itertools

import numpy as np


def _pbc3(pbc):
    pbc = np.asarray(pbc, dtype=bool)
    return np.full(3, bool(pbc)) if pbc.ndim == 0 else pbc.reshape(3)


def complete_cell(cell):
    cell = np.array(cell, dtype=float).reshape(3, 3)
    missing = np.nonzero(~cell.any(axis=1))[0]
    if len(missing) == 3:
        cell.flat[::4] = 1.0
    elif len(missing) == 2:
        i = np.nonzero(cell.any(axis=1))[0][0]
        a = cell[i]
        t = np.eye(3)[np.argmin(np.abs(a))]
        b = np.cross(a, t)
        cell[missing[0]] = b / np.linalg.norm(b)
        c = np.cross(a, cell[missing[0]])
        cell[missing[1]] = c / np.linalg.norm(c)
    elif len(missing) == 1:
        i = missing[0]
        c = np.cross(cell[i - 2], cell[i - 1])
        cell[i] = c / np.linalg.norm(c)
    return cell


def wrap_positions(positions, cell, pbc=True, center=(0.5, 0.5, 0.5),
                   pretty_translation=False, eps=1e-7):
    if not hasattr(center, '__len__'):
        center = (center,) * 3
    pbc = _pbc3(pbc)
    shift = np.asarray(center) - 0.5 - eps
    shift[np.logical_not(pbc)] = 0.0
    cell = complete_cell(np.asarray(cell))
    fractional = np.linalg.solve(cell.T, np.asarray(positions, dtype=float).T).T - shift
    for i, periodic in enumerate(pbc):
        if periodic:
            fractional[:, i] %= 1.0
            fractional[:, i] += shift[i]
    return np.dot(fractional, cell)


def find_mic(v, cell, pbc=True):
    """Minimum-image vectors and their lengths."""
    v = np.array(v, dtype=float)
    single = v.ndim == 1
    v = v.reshape(-1, 3)
    pbc = _pbc3(pbc)
    if not pbc.any():
        out = v
    else:
        cc = complete_cell(np.asarray(cell))
        frac = np.linalg.solve(cc.T, v.T).T
        frac[:, pbc] -= np.round(frac[:, pbc])
        base = frac @ cc
        off = np.array(list(itertools.product(*[(-1, 0, 1) if p else (0,) for p in pbc])),
                       dtype=float) @ cc
        if np.allclose(cc - np.diag(np.diag(cc)), 0.0):
            out = base
        else:
            cand = base[:, None, :] + off[None, :, :]
            best = np.argmin((cand ** 2).sum(-1), axis=1)
            out = cand[np.arange(len(base)), best]
    lengths = np.linalg.norm(out, axis=1)
    if single:
        return out[0], lengths[0]
    return out, lengths


def get_distances(p1, p2=None, cell=None, pbc=None):
    p1 = np.atleast_2d(p1)
    p2 = p1 if p2 is None else np.atleast_2d(p2)
    D = (p2[None, :, :] - p1[:, None, :]).reshape(-1, 3)
    if cell is not None and pbc is not None and np.any(pbc):
        D, L = find_mic(D, cell, pbc)
    else:
        L = np.linalg.norm(D, axis=1)
    return D.reshape(len(p1), len(p2), 3), L.reshape(len(p1), len(p2))
