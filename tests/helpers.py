"""Test helpers: a minimal GCMC replica with the attributes the replica drivers touch."""
from random import Random

import numpy as np

from mcpy_b200.ensembles.acceptance import Units, accept_trial


class _Selector:
    """One of insertion / deletion / displacement per trial, box cell, own RNG stream."""

    def __init__(self, seed, L, max_disp=0.3, n_moves=1):
        self.rng = Random(seed)
        self.L, self.max_disp, self.n_moves = L, max_disp, n_moves
        self.accepted = 0

    def do_trial_move(self, atoms):
        from ase import Atoms
        u = self.rng.random()
        n = len(atoms)
        if u < 1 / 3:
            p = [self.rng.random() * self.L for _ in range(3)]
            atoms += Atoms('H', positions=[p])
            return atoms, 1, 'H'
        if u < 2 / 3:
            if n <= 1:
                return False, -1, 'H'
            del atoms[self.rng.randrange(n)]
            return atoms, -1, 'H'
        if n == 0:
            return False, 0, 'X'
        i = self.rng.randrange(n)
        atoms.positions[i] += [(2 * self.rng.random() - 1) * self.max_disp for _ in range(3)]
        return atoms, 0, 'X'

    def get_volume(self):
        return self.L ** 3

    def get_exchange_count(self):
        return None

    def acceptance_counter(self):
        self.accepted += 1


class MiniGCMC:
    """The slice of ``GrandCanonicalEnsemble`` the batched / sharded drivers use."""

    def __init__(self, atoms, L, temperature, mu, seed):
        self.atoms = atoms
        self.E_old = 0.0
        self.n_atoms = len(atoms)
        self._mu = dict(mu)
        self.units = Units('LJ', temperature, {'H': 1.0})
        self.move_selector = _Selector(seed + 1, L)
        self._acc_rng = Random(seed + 2)
        self._wrap_on_accept = True
        self._step = 0
        self.volume_refreshes = 0

    def _acceptance_condition(self, dE, dN, volume, species, n_atoms):
        return accept_trial(self.units, self._mu, dE, dN, volume, species, n_atoms,
                            lambda: self._acc_rng.uniform(0.0, 1.0))

    def _species_count(self, atoms, specie):
        return atoms.get_chemical_symbols().count(specie)

    def _minimum_score(self, atoms, energy):
        score = energy
        for s, mu in self._mu.items():
            score -= mu * self._species_count(atoms, s)
        return score

    def calculate_cells_volume(self, atoms):
        self.volume_refreshes += 1


class OracleLJCalculator:
    """CPU calculator with the batched contract (test stand-in for B200Calculator)."""

    def __init__(self, rc=2.5):
        self.rc = rc

    def get_potential_energy(self, atoms):
        from oracle import cfast
        return cfast.lj_energy(atoms.positions, np.asarray(atoms.cell), atoms.pbc, 1.0, 1.0, self.rc)

    def get_potential_energies(self, atoms_list, chunk_size=None):
        if not atoms_list:
            return np.empty(0)
        return np.array([self.get_potential_energy(a) for a in atoms_list])


def make_factory(L=7.0, n0=24, base_seed=100):
    def factory(T=None, mu=None, rank=0):
        from ase import Atoms
        rng = np.random.default_rng(base_seed + rank)
        atoms = Atoms(f'H{n0}', positions=rng.uniform(0.5, L - 0.5, size=(n0, 3)), cell=[L, L, L], pbc=True)
        return MiniGCMC(atoms, L, T if T is not None else 2.0, mu if mu is not None else {'H': -2.5},
                        seed=base_seed + 10 * rank)
    return factory
