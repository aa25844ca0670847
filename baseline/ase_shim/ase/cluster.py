"""``ase.cluster.Octahedron`` stand-in: fcc octahedron with ``length`` atoms on
an edge, ``cutoff`` layers removed at each vertex by {100} planes."""
import numpy as np

from .atoms import Atoms
from .build import _LATTICE


def Octahedron(symbol, length, cutoff=0, latticeconstant=None, alloy=False):
    a = _LATTICE[symbol] if latticeconstant is None else latticeconstant
    m = length - 1
    rng = np.arange(-m, m + 1)
    i, j, k = np.meshgrid(rng, rng, rng, indexing='ij')
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    keep = ((np.abs(i) + np.abs(j) + np.abs(k)) <= m) & (((i + j + k) - m) % 2 == 0)
    keep &= (np.abs(i) <= m - cutoff) & (np.abs(j) <= m - cutoff) & (np.abs(k) <= m - cutoff)
    pos = np.stack([i[keep], j[keep], k[keep]], axis=1) * (a / 2.0)
    pos = pos - pos.min(axis=0)
    return Atoms([symbol] * len(pos), positions=pos)
