"""Insertion proposals whose ``min_insert`` rejection loop runs on the GPU.

Mirrors ``InsertionMove`` (``mcpy/moves/insertion_move.py:12-74``) and
``MoleculeInsertionMove`` (``mcpy/moves/molecule_insertion_move.py:14-89``): same constructor
arguments, same ``do_trial_move(atoms) -> (atoms | False, +1, species)`` contract (``atoms`` mutated in
place, ``False`` sentinel when the species cap is hit or no point is found in 1000 attempts), same
random streams.  The reference tests one candidate point per Python iteration with an O(N) numpy
distance pass; here candidates are drawn in batches, tested in one launch (``mcpyb_min_dist2_host``)
and the generators are then rewound and replayed up to the first accepted candidate, so a run seeded
like the reference consumes exactly the draws the reference would.

The distance test runs on a PRIVATE query engine of the move (same device as the ``engine`` it is
given, cell list cut at ``min_insert``): uploading the proposal's configuration there leaves the resident
bases of the calculator's engine -- the incremental ``get_potential_energy`` / ``get_potential_energies``
path -- untouched.
"""
from __future__ import annotations

import copy
from random import Random

import numpy as np

_MAX_INSERT_ATTEMPTS = 1000


class _Rng:
    """``mcpy.utils.RandomNumberGenerator`` (``random_number_generator.py:4-47``)."""

    def __init__(self, seed=None):
        self.random = Random(seed)

    def get_uniform(self, a=0.0, b=1.0):
        return self.random.uniform(a, b)

    def get_gaussian(self, mu=0.0, sigma=1.0):
        return self.random.gauss(mu, sigma)


def random_rotation_matrix(rng):
    """Uniform rotation from a normalised 4-Gaussian quaternion (``molecule_utils.py:79-88``)."""
    q = np.array([rng.get_gaussian() for _ in range(4)])
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def _cell_rng_state(cell):
    rng = getattr(cell, '_rng', None)
    return None if rng is None else copy.deepcopy(rng.bit_generator.state)


def _restore_cell_rng(cell, state):
    if state is not None:
        cell._rng.bit_generator.state = state


class _B200InsertionBase:
    batch = 64

    def _query_engine(self, engine):
        """The move's own engine for ``min_dist2`` queries: a pure distance search, so any pair potential
        with cutoff ``min_insert`` gives the right cell list (atoms farther away cannot fail the test)."""
        from ..engine import Engine
        if self.min_insert is None:
            return None
        if not self.min_insert > 0:
            raise ValueError(f'min_insert must be positive, got {self.min_insert}')
        q = Engine(getattr(engine, 'device', 0) if engine is not None else 0)
        q.set_lj(1.0, 1.0, rc=float(self.min_insert) * (1.0 + 1e-9))
        return q

    def _upload(self, atoms):
        self.engine.upload([atoms.positions], [np.ones(len(atoms), dtype=np.int32)],
                           np.asarray(atoms.cell, dtype=np.float64).reshape(1, 9),
                           np.asarray(atoms.pbc, dtype=bool).reshape(1, 3))

    def get_volume(self):
        return self.cell.get_volume()

    def calculate_volume(self, atoms):
        self.cell.calculate_volume(atoms)


class B200InsertionMove(_B200InsertionBase):
    def __init__(self, cell, species, seed, engine=None, min_insert=None, max_atoms=None):
        self.cell, self.species = cell, species
        self.rng = _Rng(seed=seed)
        self.min_insert, self.max_atoms = min_insert, max_atoms
        self._min_insert_sq = min_insert ** 2 if min_insert is not None else None
        self.engine = self._query_engine(engine)

    def do_trial_move(self, atoms):
        from ase import Atoms
        selected = self.rng.random.choice(self.species)
        if self.max_atoms is not None and atoms.get_chemical_symbols().count(selected) >= self.max_atoms:
            return False, 1, selected
        point = self.cell.get_random_point()
        if self._min_insert_sq is not None:
            idx = self.cell.get_atoms_specie_inside_cell(atoms, self.cell.get_species())
            if len(idx) > 0:
                mask = np.zeros(len(atoms), dtype=bool)
                mask[idx] = True
                self._upload(atoms)
                # Sequential semantics of the reference (test the current point, draw the next one
                # only after a failure), evaluated in batches: attempt a uses the a-th point drawn
                # after `state`; the generator is rewound and replayed to the accepted attempt.
                state = _cell_rng_state(self.cell)
                placed, attempts, first_point = False, 0, point
                while attempts < _MAX_INSERT_ATTEMPTS and not placed:
                    nb = min(self.batch, _MAX_INSERT_ATTEMPTS - attempts)
                    pts = [first_point] if first_point is not None else []
                    first_point = None
                    while len(pts) < nb:
                        pts.append(self.cell.get_random_point())
                    d2 = self.engine.min_dist2(np.array(pts), 0, mask)
                    ok = np.nonzero(d2 >= self._min_insert_sq)[0]
                    if len(ok):
                        first = attempts + int(ok[0])
                        point = pts[int(ok[0])]
                        placed = True
                        _restore_cell_rng(self.cell, state)
                        for _ in range(first):
                            self.cell.get_random_point()
                        break
                    attempts += len(pts)
                if not placed:
                    # the reference has drawn one new point after each of its 1000 failures
                    _restore_cell_rng(self.cell, state)
                    for _ in range(_MAX_INSERT_ATTEMPTS):
                        self.cell.get_random_point()
                    return False, 1, selected
        atoms += Atoms(selected, positions=[point])
        if 'molecule_id' in atoms.arrays:
            atoms.arrays['molecule_id'][-1] = -1
        return atoms, 1, selected


class B200MoleculeInsertionMove(_B200InsertionBase):
    """Rigid-molecule insertion; the (atoms x fragment-atoms) distance matrix of the reference
    (``molecule_insertion_move.py:61-72``) becomes one min-distance query per fragment atom."""

    def __init__(self, cell, molecule, name, seed, engine=None, min_insert=None, max_molecules=None,
                 count_molecules=None):
        self.cell, self.name = cell, name
        self.species = [name]
        self.rng = _Rng(seed=seed)
        self.molecule = molecule.copy()
        self.molecule.positions -= self.molecule.get_center_of_mass()
        self._template_symbols = sorted(self.molecule.get_chemical_symbols())
        self.min_insert, self.max_molecules = min_insert, max_molecules
        self._min_insert_sq = min_insert ** 2 if min_insert is not None else None
        self.last_exchange_count = None
        # callable(atoms, template_symbols, cell) -> list of member index arrays; the default is the cached
        # equivalent of find_molecules(atoms, template, cell) (molecule_insertion_move.py:44): composition match
        # AND centre of mass exchangeable with the cell -- the count the deletion move uses too
        self._count = count_molecules
        self._index = None
        self.engine = self._query_engine(engine)

    def _n_molecules(self, atoms):
        if self._count is not None:
            return len(self._count(atoms, self._template_symbols, self.cell))
        if self._index is None:
            from .molecules import MoleculeIndex
            self._index = MoleculeIndex(self._template_symbols)
        return len(self._index.groups(atoms, self.cell))

    def do_trial_move(self, atoms):
        n_mol = self._n_molecules(atoms)
        if self.max_molecules is not None and n_mol >= self.max_molecules:
            return False, 1, self.name
        self.last_exchange_count = n_mol
        point = self.cell.get_random_point()
        rotation = random_rotation_matrix(self.rng)
        k = len(self.molecule)
        if self._min_insert_sq is not None:
            idx = self.cell.get_atoms_specie_inside_cell(atoms, self.cell.get_species())
            if len(idx) > 0:
                mask = np.zeros(len(atoms), dtype=bool)
                mask[idx] = True
                self._upload(atoms)
                placed = False
                # sequential semantics of the reference, evaluated in batches; both generators are
                # rewound and replayed to the accepted attempt
                attempts = 0
                cell_state = _cell_rng_state(self.cell)
                py_state = self.rng.random.getstate()
                first_pair = (point, rotation)
                while attempts < _MAX_INSERT_ATTEMPTS and not placed:
                    nb = min(self.batch, _MAX_INSERT_ATTEMPTS - attempts)
                    pairs = [first_pair] if first_pair is not None else []
                    first_pair = None
                    while len(pairs) < nb:
                        pairs.append((self.cell.get_random_point(), random_rotation_matrix(self.rng)))
                    frags = np.concatenate([self.molecule.positions @ R.T + p for p, R in pairs])
                    d2 = self.engine.min_dist2(frags, 0, mask).reshape(len(pairs), k).min(axis=1)
                    ok = np.nonzero(d2 >= self._min_insert_sq)[0]
                    if len(ok):
                        first = attempts + int(ok[0])
                        point, rotation = pairs[int(ok[0])]
                        placed = True
                        # replay: attempt 0 was drawn before the loop; attempts 1..first follow
                        _restore_cell_rng(self.cell, cell_state)
                        self.rng.random.setstate(py_state)
                        for _ in range(first):
                            self.cell.get_random_point()
                            random_rotation_matrix(self.rng)
                        break
                    attempts += len(pairs)
                if not placed:
                    # the reference has drawn one (point, rotation) after each of its 1000 failures
                    _restore_cell_rng(self.cell, cell_state)
                    self.rng.random.setstate(py_state)
                    for _ in range(_MAX_INSERT_ATTEMPTS):
                        self.cell.get_random_point()
                        random_rotation_matrix(self.rng)
                    return False, 1, self.name
        fragment = self.molecule.copy()
        fragment.positions = fragment.positions @ rotation.T + point
        ids = atoms.arrays.get('molecule_id')
        if ids is None:
            atoms.new_array('molecule_id', np.full(len(atoms), -1, dtype=int))
            next_id = 0
        else:
            next_id = int(ids.max()) + 1 if (ids >= 0).any() else 0
        fragment.new_array('molecule_id', np.full(len(fragment), next_id, dtype=int))
        atoms += fragment
        return atoms, 1, self.name
