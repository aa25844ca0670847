"""``FixAtoms`` stand-in with ASE's index remapping on atom deletion."""
import numpy as np


class FixAtoms:
    def __init__(self, indices=None, mask=None):
        if mask is not None:
            indices = np.nonzero(np.asarray(mask, bool))[0]
        self.index = np.unique(np.asarray(indices if indices is not None else [], dtype=int))

    def get_indices(self):
        return self.index.copy()

    def copy(self):
        return FixAtoms(indices=self.index.copy())

    def delete_atoms(self, indices, natoms):
        i = np.zeros(natoms, int) - 1
        new = np.delete(np.arange(natoms), indices)
        i[new] = np.arange(len(new))
        index = i[self.index]
        self.index = index[index >= 0]
        if len(self.index) == 0:
            return None
        return self

    def adjust_positions(self, atoms, new):
        new[self.index] = atoms.positions[self.index]

    def adjust_forces(self, atoms, forces):
        forces[self.index] = 0.0

    def __repr__(self):
        return f'FixAtoms(indices={self.index.tolist()})'
