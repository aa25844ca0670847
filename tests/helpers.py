"""Test helpers: CPU stand-in calculators and a factory of UNMODIFIED reference GCMC replicas."""
import numpy as np


def de_tol(dE_ref, E_scale):
    """Tolerance on a trial-move dE: 1e-10 relative to dE itself (north_star), plus the rounding of the ORACLE's
    difference of two totals of size ``E_scale`` (each total is the fp64 rounding of an exact long-double sum:
    <= 1.2e-16 |E| per total; 1e-13 leaves room for the numpy restatements)."""
    return 1e-10 * max(1.0, abs(dE_ref)) + 1e-13 * abs(E_scale)


class OracleLJCalculator:
    """CPU calculator with the batched contract (test stand-in for B200Calculator)."""

    def __init__(self, rc=2.5, sigma=1.0, epsilon=1.0):
        self.rc, self.sigma, self.epsilon = rc, sigma, epsilon

    def get_potential_energy(self, atoms):
        from oracle import cfast
        return cfast.lj_energy(atoms.positions, np.asarray(atoms.cell), atoms.pbc, self.sigma, self.epsilon, self.rc)

    def get_potential_energies(self, atoms_list, chunk_size=None):
        if not atoms_list:
            return np.empty(0)
        return np.array([self.get_potential_energy(a) for a in atoms_list])


class OracleEMTCalculator:
    """The EMT oracle (plain C) behind the calculator interface."""

    def get_potential_energy(self, atoms):
        from oracle import cfast
        return cfast.emt_energy(atoms.positions, atoms.numbers, np.asarray(atoms.cell), atoms.pbc)

    def get_potential_energies(self, atoms_list, chunk_size=None):
        if not atoms_list:
            return np.empty(0)
        return np.array([self.get_potential_energy(a) for a in atoms_list])


def make_factory(calc, L=7.0, n0=24, base_seed=100, fix=(), n_moves=1, extra_mu=None, origin=True):
    """``gcmc_factory(T=.., rank=..)`` / ``(mu=.., rank=..)`` of the replica drivers: a fresh reference
    ``GrandCanonicalEnsemble`` (own atoms, cell, moves, selector, seeds from ``rank``) in LJ units.
    ``fix``: indices held by an ``ase.constraints.FixAtoms``; ``extra_mu``: ``{rank: mu dict}`` overrides
    (ladders whose replicas carry different chemical-potential key sets)."""
    def factory(T=None, mu=None, rank=0):
        from ase import Atoms
        from ase.constraints import FixAtoms
        from mcpy.cell import Cell
        from mcpy.ensembles import GrandCanonicalEnsemble
        from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
        rng = np.random.default_rng(base_seed + rank)
        atoms = Atoms(f'H{n0}', positions=rng.uniform(0.5, L - 0.5, size=(n0, 3)), cell=[L, L, L], pbc=True)
        if len(fix):
            atoms.set_constraint(FixAtoms(indices=list(fix)))
        if origin:
            atoms.info['origin'] = rank
        cell = Cell(atoms, seed=300 + rank)
        moves = [InsertionMove(cell, species=['H'], seed=11 + rank), DeletionMove(cell, species=['H'], seed=12 + rank),
                 DisplacementMove(species=['H'], seed=13 + rank, max_displacement=0.3)]
        ms = MoveSelector([1, 1, 1], moves, seed=14 + rank, n_moves=n_moves)
        mu_r = (extra_mu or {}).get(rank, mu if mu is not None else {'H': -2.5})
        return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=calc,
                                      mu=mu_r, species=['H'], temperature=T if T is not None else 2.0,
                                      move_selector=ms, random_seed=40 + rank, traj_file=None, outfile=None)
    return factory
