"""Minimal ``ase.cell.Cell``: a 3x3 array with a few helpers."""
import numpy as np


class Cell:
    def __init__(self, array=None):
        self.array = np.zeros((3, 3)) if array is None else np.array(array, dtype=float).reshape(3, 3)

    @classmethod
    def new(cls, cell=None):
        if cell is None:
            return cls(np.zeros((3, 3)))
        if isinstance(cell, Cell):
            return cls(cell.array.copy())
        cell = np.array(cell, dtype=float)
        if cell.shape == (3,):
            return cls(np.diag(cell))
        if cell.shape == (3, 3):
            return cls(cell)
        raise ValueError('cell must be 3 lengths or a 3x3 matrix (shim)')

    @property
    def volume(self):
        return float(abs(np.linalg.det(self.array)))

    def lengths(self):
        return np.linalg.norm(self.array, axis=1)

    def copy(self):
        return Cell(self.array.copy())

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)

    def __getitem__(self, item):
        return self.array[item]

    def __setitem__(self, item, value):
        self.array[item] = value

    def __len__(self):
        return 3

    def __iter__(self):
        return iter(self.array)

    def any(self):
        return bool(self.array.any())

    def __repr__(self):
        return f'Cell({self.array.tolist()})'
