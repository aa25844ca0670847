"""Minimal ``ase.Atoms`` stand-in (TEST-ONLY; see package docstring)."""
import copy as _copy
import re

import numpy as np

from .cell import Cell
from .data import atomic_masses, atomic_numbers, chemical_symbols
from .geometry import wrap_positions


def _parse_symbols(symbols):
    if symbols is None:
        return []
    if isinstance(symbols, str):
        out = []
        for sym, count in re.findall(r'([A-Z][a-z]?)(\d*)', symbols):
            out.extend([atomic_numbers[sym]] * (int(count) if count else 1))
        return out
    out = []
    for s in symbols:
        out.append(atomic_numbers[s] if isinstance(s, str) else int(s))
    return out


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None,
                 pbc=None, calculator=None, info=None, constraint=None):
        if isinstance(symbols, Atoms):
            other = symbols
            numbers = other.numbers.copy()
            positions = other.positions.copy()
            cell = other.cell if cell is None else cell
            pbc = other.pbc if pbc is None else pbc
            symbols = None
        if numbers is None:
            numbers = _parse_symbols(symbols)
        numbers = np.array(numbers, dtype=int).reshape(-1)
        n = len(numbers)
        if positions is None:
            positions = np.zeros((n, 3))
        positions = np.array(positions, dtype=float).reshape(-1, 3)
        if len(positions) != n:
            raise ValueError('positions and symbols have different lengths (shim)')
        self.arrays = {'numbers': numbers, 'positions': positions}
        self._cellobj = Cell.new(cell)
        self._pbc = np.zeros(3, dtype=bool)
        if pbc is not None:
            self.pbc = pbc
        self.calc = calculator
        self.info = dict(info or {})
        self._constraints = []
        if constraint is not None:
            self.set_constraint(constraint)

    # -- basic properties --------------------------------------------------
    @property
    def positions(self):
        return self.arrays['positions']

    @positions.setter
    def positions(self, value):
        self.arrays['positions'][:] = value

    @property
    def numbers(self):
        return self.arrays['numbers']

    @numbers.setter
    def numbers(self, value):
        self.arrays['numbers'][:] = value

    @property
    def cell(self):
        return self._cellobj

    @cell.setter
    def cell(self, value):
        self._cellobj = Cell.new(value)

    @property
    def pbc(self):
        return self._pbc

    @pbc.setter
    def pbc(self, value):
        self._pbc[:] = value

    @property
    def constraints(self):
        return self._constraints

    @constraints.setter
    def constraints(self, value):
        self.set_constraint(value)

    def set_constraint(self, constraint=None):
        if constraint is None:
            self._constraints = []
        elif isinstance(constraint, (list, tuple)):
            self._constraints = list(constraint)
        else:
            self._constraints = [constraint]

    def set_cell(self, cell, scale_atoms=False):
        self._cellobj = Cell.new(cell)

    def set_pbc(self, pbc):
        self.pbc = pbc

    def get_cell(self, complete=False):
        return self._cellobj.copy()

    def get_pbc(self):
        return self._pbc.copy()

    def get_positions(self, wrap=False):
        return self.positions.copy()

    def set_positions(self, newpositions, apply_constraint=True):
        self.arrays['positions'][:] = newpositions

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [chemical_symbols[z] for z in self.numbers]

    def get_chemical_formula(self):
        syms = self.get_chemical_symbols()
        return ''.join(f'{s}{syms.count(s) if syms.count(s) > 1 else ""}'
                       for s in sorted(set(syms)))

    def get_masses(self):
        if 'masses' in self.arrays:
            return self.arrays['masses'].copy()
        return atomic_masses[self.numbers]

    def get_center_of_mass(self, scaled=False):
        m = self.get_masses()
        return m @ self.positions / m.sum()

    def get_volume(self):
        return self._cellobj.volume

    @property
    def symbols(self):
        return _Symbols(self.arrays['numbers'])

    @symbols.setter
    def symbols(self, value):
        self.arrays['numbers'][:] = _parse_symbols(value)

    def get_distance(self, a0, a1, mic=False, vector=False):
        from .geometry import find_mic
        d = self.positions[a1] - self.positions[a0]
        if mic:
            d = find_mic(d, self._cellobj.array, self._pbc)[0]
        return d if vector else float(np.linalg.norm(d))

    def get_all_distances(self, mic=False, vector=False):
        from .geometry import get_distances
        D, L = get_distances(self.positions, cell=self._cellobj.array if mic else None,
                             pbc=self._pbc if mic else None)
        return D if vector else L

    def translate(self, displacement):
        self.arrays['positions'] += np.array(displacement)

    def center(self, vacuum=None, axis=(0, 1, 2), about=None):
        if about is not None:
            self.translate(np.asarray(about) - self.positions.mean(axis=0))
            return
        p = self.positions
        if vacuum is not None:
            ext = p.max(axis=0) - p.min(axis=0)
            cell = np.diag(ext + 2 * vacuum)
            self._cellobj = Cell.new(cell)
        c = self._cellobj.array.sum(axis=0) / 2
        self.translate(c - (p.max(axis=0) + p.min(axis=0)) / 2)

    def wrap(self, **wrap_kw):
        if 'pbc' not in wrap_kw:
            wrap_kw['pbc'] = self.pbc
        self.positions[:] = wrap_positions(self.positions, self._cellobj.array, **wrap_kw)

    def new_array(self, name, a, dtype=None, shape=None):
        a = np.array(a, dtype=dtype, order='C')
        if name in self.arrays:
            raise RuntimeError(f'Array {name} already present')
        if len(a) != len(self):
            raise ValueError('Array has wrong length')
        self.arrays[name] = a

    def get_array(self, name, copy=True):
        return self.arrays[name].copy() if copy else self.arrays[name]

    def set_array(self, name, a, dtype=None, shape=None):
        if a is None:
            self.arrays.pop(name, None)
            return
        a = np.array(a, dtype=dtype)
        if name in self.arrays and self.arrays[name].shape == a.shape:
            self.arrays[name][:] = a
        else:
            self.arrays[name] = a

    def has(self, name):
        return name in self.arrays

    # -- container protocol ------------------------------------------------
    def __len__(self):
        return len(self.arrays['positions'])

    def copy(self):
        new = Atoms.__new__(Atoms)
        new.arrays = {k: v.copy() for k, v in self.arrays.items()}
        new._cellobj = self._cellobj.copy()
        new._pbc = self._pbc.copy()
        new.calc = None
        new.info = _copy.deepcopy(self.info)
        new._constraints = [c.copy() for c in self._constraints]
        return new

    def extend(self, other):
        if not isinstance(other, Atoms):
            other = Atoms(other)
        n1, n2 = len(self), len(other)
        for name, a1 in list(self.arrays.items()):
            a = np.zeros((n1 + n2,) + a1.shape[1:], a1.dtype)
            a[:n1] = a1
            if name == 'masses':
                a2 = other.get_masses()
            else:
                a2 = other.arrays.get(name)
            if a2 is not None:
                a[n1:] = a2
            self.arrays[name] = a
        for name, a2 in other.arrays.items():
            if name in self.arrays:
                continue
            a = np.empty((n1 + n2,) + a2.shape[1:], a2.dtype)
            a[n1:] = a2
            if name == 'masses':
                a[:n1] = self.get_masses()[:n1]
            else:
                a[:n1] = 0
            self.arrays[name] = a

    def __iadd__(self, other):
        self.extend(other)
        return self

    def __add__(self, other):
        new = self.copy()
        new += other
        return new

    def append(self, atom):
        self.extend(Atoms([atom.symbol], positions=[atom.position]))

    def pop(self, i=-1):
        atom = Atom(self[i].symbol, position=self[i].position.copy())
        del self[i]
        return atom

    def __delitem__(self, i):
        if isinstance(i, (list, tuple)) and len(i) > 0:
            i = np.array(i)
        if len(self._constraints) > 0:
            n = len(self)
            idx = np.arange(n)[i]
            if isinstance(idx, (int, np.integer)):
                idx = [idx]
            new_constraints = []
            for con in self._constraints:
                con = con.delete_atoms(idx, n) if hasattr(con, 'delete_atoms') else con
                if con is not None:
                    new_constraints.append(con)
            self._constraints = new_constraints
        mask = np.ones(len(self), bool)
        mask[i] = False
        for name, a in self.arrays.items():
            self.arrays[name] = a[mask]

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            n = len(self)
            if i < -n or i >= n:
                raise IndexError('Index out of range.')
            return _Atom(self, int(i) % n)
        if isinstance(i, (list, tuple)) and len(i) > 0 and isinstance(i[0], (bool, np.bool_)):
            i = np.array(i, dtype=bool)
        elif isinstance(i, (list, tuple)):
            i = np.array(i, dtype=int)
        new = Atoms.__new__(Atoms)
        new.arrays = {k: v[i].copy() for k, v in self.arrays.items()}
        new._cellobj = self._cellobj.copy()
        new._pbc = self._pbc.copy()
        new.calc = None
        new.info = _copy.deepcopy(self.info)
        new._constraints = []
        return new

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __imul__(self, m):
        if isinstance(m, int):
            m = (m, m, m)
        n = len(self)
        cellarr = self._cellobj.array
        reps = [(i, j, k) for i in range(m[0]) for j in range(m[1]) for k in range(m[2])]
        for name, a in self.arrays.items():
            self.arrays[name] = np.tile(a, (len(reps),) + (1,) * (a.ndim - 1))
        pos = self.arrays['positions']
        for r, (i, j, k) in enumerate(reps):
            pos[r * n:(r + 1) * n] += np.array([i, j, k], float) @ cellarr
        self._cellobj = Cell(np.array([m[c] * cellarr[c] for c in range(3)]))
        return self

    def __mul__(self, m):
        new = self.copy()
        new *= m
        return new

    repeat = __mul__

    # -- calculator protocol -----------------------------------------------
    def get_potential_energy(self, force_consistent=False, apply_constraint=True):
        if self.calc is None:
            raise RuntimeError('Atoms object has no calculator.')
        return self.calc.get_potential_energy(self)

    def get_forces(self, apply_constraint=True, md=False):
        if self.calc is None:
            raise RuntimeError('Atoms object has no calculator.')
        forces = np.array(self.calc.get_forces(self), dtype=float)
        if apply_constraint:
            for c in self.constraints:
                c.adjust_forces(self, forces)
        return forces

    def __repr__(self):
        return (f'Atoms(symbols={self.get_chemical_formula()!r}, pbc={self._pbc.tolist()}, '
                f'cell={self._cellobj.array.tolist()})')


class _Atom:
    def __init__(self, atoms, index):
        self.atoms, self.index = atoms, index

    @property
    def symbol(self):
        return chemical_symbols[self.atoms.numbers[self.index]]

    @property
    def number(self):
        return int(self.atoms.numbers[self.index])

    @property
    def position(self):
        return self.atoms.positions[self.index]

    @position.setter
    def position(self, value):
        self.atoms.positions[self.index] = value

    @property
    def mass(self):
        return float(self.atoms.get_masses()[self.index])


class Atom:
    """Free-standing atom (``ase.Atom``): symbol + position."""

    def __init__(self, symbol='X', position=(0, 0, 0)):
        self.symbol = symbol if isinstance(symbol, str) else chemical_symbols[symbol]
        self.position = np.array(position, dtype=float)

    @property
    def number(self):
        return atomic_numbers[self.symbol]


class _Symbols:
    """View on the atomic numbers with a list-of-symbols face (``atoms.symbols``)."""

    def __init__(self, numbers):
        self.numbers = numbers

    def __len__(self):
        return len(self.numbers)

    def __iter__(self):
        return (chemical_symbols[z] for z in self.numbers)

    def __getitem__(self, i):
        z = self.numbers[i]
        if np.ndim(z) == 0:
            return chemical_symbols[int(z)]
        return _Symbols(z)

    def __setitem__(self, i, value):
        if isinstance(value, str):
            self.numbers[i] = atomic_numbers[value]
        else:
            self.numbers[i] = [atomic_numbers[v] if isinstance(v, str) else v for v in value]

    def count(self, sym):
        return int(np.count_nonzero(self.numbers == atomic_numbers[sym]))

    def __eq__(self, other):
        return list(self) == list(other)

    def __str__(self):
        return ''.join(self)
