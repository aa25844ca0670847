"""Rigid-molecule deletion and displacement proposals with O(N) bookkeeping.

Mirrors ``MoleculeDeletionMove`` (``mcpy/moves/molecule_deletion_move.py:10-43``) and
``MoleculeDisplacementMove`` (``mcpy/moves/molecule_displacement_move.py:11-104``): same constructor
arguments, same ``do_trial_move(atoms) -> (atoms | False, dN, name)`` contract (``atoms`` mutated in
place), same random streams and the same floating-point results -- a run seeded like the reference
proposes the same moves bit for bit (``tests/test_molecule_moves_cpu.py`` replays a trace recorded from
the unmodified classes).

What changes is the cost of finding the candidates.  The reference's ``find_molecules``
(``molecule_utils.py:51-76``) scans the whole ``molecule_id`` array once per molecule and evaluates the
centre-of-mass predicate of every molecule on every trial: O(M N) numpy calls plus M minimum-image
reductions -- milliseconds per trial for 200 CO on a 10 000-atom particle, two orders of magnitude more
than the energy evaluation behind it.  ``MoleculeIndex`` groups the atoms with one
stable argsort, and remembers each molecule's centre of mass and predicate result together with the
positions they were computed from, so a trial re-evaluates only the molecules that moved.
"""
from __future__ import annotations

import numpy as np

from .insertion import _Rng, random_rotation_matrix


def exchangeable_predicate(cell):
    """``molecule_utils.exchangeable_predicate`` (``molecule_utils.py:40-48``)."""
    return getattr(cell, 'is_point_exchangeable', None) or cell.is_point_inside


def molecule_com(atoms, members):
    """``molecule_utils.molecule_com`` (``molecule_utils.py:18-37``), operation for operation."""
    from ase.geometry import find_mic, wrap_positions
    pos = atoms.positions[members]
    ref = pos[0]
    diff = pos - ref
    use_mic = bool(np.any(np.asarray(atoms.pbc)))
    if use_mic:
        diff = find_mic(diff, atoms.cell, pbc=atoms.pbc)[0]
    masses = atoms.get_masses()[members]
    com = ref + (diff * masses[:, None]).sum(axis=0) / masses.sum()
    if use_mic:
        com = wrap_positions([com], atoms.cell, pbc=atoms.pbc)[0]
    return com


class MoleculeIndex:
    """Candidate molecules of one template, cached across trials.

    ``groups(atoms, cell)`` returns what ``find_molecules(atoms, template_symbols, cell)`` returns: the
    member index arrays (ascending atom index) of the molecules whose composition matches the template
    and whose centre of mass the cell counts as exchangeable, in ascending ``molecule_id`` order.
    """

    def __init__(self, template_symbols):
        from ase.data import atomic_numbers
        self.template_symbols = list(template_symbols)
        self._template_numbers = np.sort(np.array([atomic_numbers[s] for s in template_symbols]))
        # last evaluation: molecule ids (ascending), member positions (M, k, 3), centres of mass, predicate
        self._mids = np.zeros(0, dtype=np.int64)
        self._pos = np.zeros((0, len(template_symbols), 3))
        self._com = np.zeros((0, 3))
        self._flag = np.zeros(0, dtype=bool)
        self._box = None
        self.recomputed = 0         # predicate evaluations so far (for the tests)

    def _grouping(self, atoms):
        """(molecule ids ascending, members (M, k) ascending atom index) of the template's composition."""
        k = len(self._template_numbers)
        empty = (np.zeros(0, dtype=np.int64), np.zeros((0, k), dtype=np.int64))
        ids = atoms.arrays.get('molecule_id')
        if ids is None:
            return empty
        ids = np.asarray(ids)
        order = np.argsort(ids, kind='stable')              # members of a molecule in ascending atom index
        sid = ids[order]
        first = np.searchsorted(sid, 0, side='left')        # skip the -1 (not in a molecule) block
        order, sid = order[first:], sid[first:]
        if len(sid) == 0:
            return empty
        starts = np.concatenate([[0], np.flatnonzero(np.diff(sid)) + 1])
        sizes = np.diff(np.concatenate([starts, [len(sid)]]))
        starts = starts[sizes == k]                          # molecules of the template's size
        if len(starts) == 0:
            return empty
        members = order[starts[:, None] + np.arange(k)[None, :]]
        match = np.all(np.sort(np.asarray(atoms.numbers)[members], axis=1) == self._template_numbers[None, :], axis=1)
        return sid[starts][match].astype(np.int64), members[match]

    def groups(self, atoms, cell=None, with_com=False):
        mids, members = self._grouping(atoms)
        if cell is None:
            return list(members)
        predicate = exchangeable_predicate(cell)
        box = (np.asarray(atoms.cell, dtype=float).tobytes(), np.asarray(atoms.pbc, dtype=bool).tobytes(), id(cell))
        if box != self._box:
            self._mids = np.zeros(0, dtype=np.int64)
            self._box = box
        pos = atoms.positions[members]                       # (M, k, 3)
        com = np.zeros((len(mids), 3))
        flag = np.zeros(len(mids), dtype=bool)
        # reuse what was computed from bit-identical member positions of the same molecule id
        reuse = np.zeros(len(mids), dtype=bool)
        if len(self._mids) and len(mids):
            at = np.searchsorted(self._mids, mids)
            at[at >= len(self._mids)] = 0
            same = self._mids[at] == mids
            same[same] = np.all(self._pos[at[same]] == pos[same], axis=(1, 2))
            reuse = same
            com[reuse] = self._com[at[reuse]]
            flag[reuse] = self._flag[at[reuse]]
        for i in np.flatnonzero(~reuse):
            com[i] = molecule_com(atoms, members[i])
            flag[i] = bool(predicate(com[i]))
            self.recomputed += 1
        self._mids, self._pos, self._com, self._flag = mids, pos, com, flag
        out = list(members[flag])
        return (out, list(com[flag])) if with_com else out


class _B200MoleculeMoveBase:
    def __init__(self, cell, molecule, name, seed):
        self.cell, self.name = cell, name
        self.species = [name]
        self.rng = _Rng(seed=seed)
        self._template_symbols = sorted(molecule.get_chemical_symbols())
        self.index = MoleculeIndex(self._template_symbols)

    def get_volume(self):
        return self.cell.get_volume()

    def calculate_volume(self, atoms):
        self.cell.calculate_volume(atoms)


class B200MoleculeDeletionMove(_B200MoleculeMoveBase):
    """``MoleculeDeletionMove`` (``molecule_deletion_move.py:10-43``)."""

    def __init__(self, cell, molecule, name, seed, min_molecules=None):
        super().__init__(cell, molecule, name, seed)
        self.min_molecules = min_molecules
        self.last_exchange_count = None

    def do_trial_move(self, atoms):
        groups = self.index.groups(atoms, self.cell)
        if len(groups) == 0:
            return False, -1, self.name
        if self.min_molecules is not None and len(groups) <= self.min_molecules:
            return False, -1, self.name
        self.last_exchange_count = len(groups)
        members = self.rng.random.choice(groups)
        del atoms[members]
        return atoms, -1, self.name


class B200MoleculeDisplacementMove(_B200MoleculeMoveBase):
    """``MoleculeDisplacementMove`` (``molecule_displacement_move.py:11-104``)."""

    def __init__(self, cell, molecule, name, seed, max_displacement=0.5, max_angle=None):
        super().__init__(cell, molecule, name, seed)
        self.max_displacement = max_displacement
        self.max_angle = max_angle

    def do_trial_move(self, atoms):
        from ase.geometry import find_mic, wrap_positions
        groups, coms = self.index.groups(atoms, self.cell, with_com=True)
        if len(groups) == 0:
            return False, 0, self.name
        pick = self.rng.random.choice(range(len(groups)))     # consumes the draws of choice(groups)
        members, com = groups[pick], coms[pick]
        rsq = 1.1
        while rsq > 1.0:
            rx = 2.0 * self.rng.get_uniform() - 1.0
            ry = 2.0 * self.rng.get_uniform() - 1.0
            rz = 2.0 * self.rng.get_uniform() - 1.0
            rsq = rx * rx + ry * ry + rz * rz
        shift = np.array([rx, ry, rz]) * self.max_displacement
        new_com = com + shift
        periodic = bool(np.any(np.asarray(atoms.pbc)))
        if periodic:
            new_com = wrap_positions([new_com], atoms.cell, pbc=atoms.pbc)[0]
        if not exchangeable_predicate(self.cell)(new_com):
            return False, 0, self.name
        rotation = random_rotation_matrix(self.rng) if self.max_angle is None else self._capped_rotation()
        offsets = atoms.positions[members] - com
        if periodic:
            offsets = find_mic(offsets, atoms.cell, pbc=atoms.pbc)[0]
        atoms.positions[members] = com + shift + offsets @ rotation.T
        return atoms, 0, self.name

    def _capped_rotation(self):
        while True:
            axis = np.array([self.rng.get_gaussian() for _ in range(3)])
            norm = np.linalg.norm(axis)
            if norm > 1e-12:
                break
        axis /= norm
        angle = self.rng.get_uniform() * self.max_angle
        w = np.cos(angle / 2.0)
        x, y, z = np.sin(angle / 2.0) * axis
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ])
