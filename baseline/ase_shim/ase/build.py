"""``bulk`` (fcc/sc cubic only) and ``molecule`` (a few diatomics/H2O)."""
import numpy as np

from .atoms import Atoms

_LATTICE = {'Cu': 3.61, 'Ag': 4.09, 'Au': 4.08, 'Ni': 3.52, 'Pd': 3.89, 'Pt': 3.92, 'Al': 4.05}
_MOLECULES = {
    'CO': (['O', 'C'], [[0.0, 0.0, 0.484676], [0.0, 0.0, -0.646514]]),   # d = 1.13119 (g2)
    'O2': (['O', 'O'], [[0.0, 0.0, 0.622978], [0.0, 0.0, -0.622978]]),
    'N2': (['N', 'N'], [[0.0, 0.0, 0.56499], [0.0, 0.0, -0.56499]]),
    'H2': (['H', 'H'], [[0.0, 0.0, 0.368583], [0.0, 0.0, -0.368583]]),
    'H2O': (['O', 'H', 'H'], [[0.0, 0.0, 0.119262], [0.0, 0.763239, -0.477047],
                              [0.0, -0.763239, -0.477047]]),
}


def bulk(name, crystalstructure='fcc', a=None, cubic=False, **kwargs):
    a = _LATTICE[name] if a is None else a
    if crystalstructure != 'fcc':
        raise NotImplementedError('shim bulk(): fcc only')
    if cubic:
        pos = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
        return Atoms([name] * 4, positions=pos, cell=[a, a, a], pbc=True)
    cell = np.array([[0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    return Atoms([name], positions=[[0, 0, 0]], cell=cell, pbc=True)


def molecule(name, **kwargs):
    syms, pos = _MOLECULES[name]
    return Atoms(syms, positions=pos)


def fcc111(symbol, size, a=None, vacuum=None, orthogonal=False, periodic=False):
    """Hexagonal fcc(111) slab, ABC stacking, ``vacuum`` on both sides in z."""
    if orthogonal:
        raise NotImplementedError('shim fcc111(): hexagonal cell only')
    a = _LATTICE[symbol] if a is None else a
    nx, ny, nz = size
    d = a / np.sqrt(2.0)
    a1 = np.array([d, 0.0, 0.0])
    a2 = np.array([d / 2.0, d * np.sqrt(3.0) / 2.0, 0.0])
    dz = a / np.sqrt(3.0)
    pos = []
    for l in range(nz):
        shift = ((nz - 1 - l) % 3) / 3.0
        for j in range(ny):
            for i in range(nx):
                pos.append((i + shift) * a1 + (j + shift) * a2 + np.array([0, 0, l * dz]))
    pos = np.array(pos)
    height = (nz - 1) * dz
    cell = np.array([nx * a1, ny * a2, [0.0, 0.0, height]])
    if vacuum is not None:
        cell[2, 2] = height + 2 * vacuum
        pos[:, 2] += vacuum
    elif cell[2, 2] == 0.0:
        cell[2, 2] = 0.0
    atoms = Atoms([symbol] * len(pos), positions=pos, cell=cell,
                  pbc=True if periodic else (True, True, False))
    atoms.wrap()
    return atoms
