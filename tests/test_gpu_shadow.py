"""GPU: the UNMODIFIED reference drivers, run twice from the same seeds.

Once with the reference's own cells and moves over the CPU oracle calculator, once with the B200
pieces plugged in behind the same interfaces -- ``B200Calculator`` as the calculator,
``B200SphericalCell`` / ``B200CustomCell`` (sample points regenerated on the device) as the cells,
the GPU-backed insertion moves and the cached molecule moves.  The drivers
(``GrandCanonicalEnsemble``, ``MoveSelector``, the atomic moves, ``BatchedReplicaExchange``,
``CanonicalEnsemble``) are the reference's code (``baseline/_ref``), never a mirror.

Asserted per trial: the same accept / reject decision, the same particle numbers, species and
de Broglie counts, the same free volume to the bit, and ``dE`` within
``1e-10 max(1, |dE|) + 1e-13 |E|`` (the second term is the rounding of a difference of two
totals of size ``|E|`` on the oracle side).  Pattern: the reference's own harness
``tests/test_calculator_gcmc_equivalence.py:70-144``.

BASELINE.json configs: C1 argon LJ (S1), C2 2 000-atom LJ (S2, repulsive and attractive cutoffs),
C3 Ag octahedron + O with EMT in a SphericalCell, C4 64 x 2 000-atom replica exchange, C5 CuPd + CO
rigid-molecule moves with EMT; plus a slab in a CustomCell and CanonicalEnsemble's relax loop.
"""
import numpy as np
import pytest

from helpers import OracleEMTCalculator, OracleLJCalculator
from mcpy_b200.utils import synthetic

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ harness
class Logged:
    """Calculator wrapper recording every energy it hands to the driver."""

    def __init__(self, inner):
        self.inner = inner
        self.energies = []

    def get_potential_energy(self, atoms):
        e = float(self.inner.get_potential_energy(atoms))
        self.energies.append(e)
        return e

    def get_potential_energies(self, atoms_list, chunk_size=None):
        e = np.asarray(self.inner.get_potential_energies(atoms_list), dtype=float)
        self.energies.extend(e.tolist())
        return e


def trace_acceptance(gcmc):
    """Record (E_old, dE, dN, volume, species, n_exchange, accepted) of every scored trial."""
    trace = []
    inner = gcmc._acceptance_condition

    def logged(dE, dN, volume, species, n):
        ok = inner(dE, dN, volume, species, n)
        trace.append((float(gcmc.E_old), float(dE), int(dN), float(volume), str(species), int(n), bool(ok)))
        return ok
    gcmc._acceptance_condition = logged
    return trace


def assert_same_traces(ref, b200, min_trials, min_accepts=1):
    assert len(ref) == len(b200) >= min_trials, (len(ref), len(b200))
    worst = 0.0
    for k, (a, b) in enumerate(zip(ref, b200)):
        assert a[2:] == b[2:], f'trial {k}: reference {a} vs B200 {b}'       # dN, volume, species, n, decision
        tol = 1e-10 * max(1.0, abs(a[1])) + 1e-13 * abs(a[0])
        assert abs(a[1] - b[1]) <= tol, f'trial {k}: dE {a[1]!r} vs {b[1]!r} (tol {tol:.3g})'
        worst = max(worst, abs(a[1] - b[1]) / max(1.0, abs(a[1])))
    assert sum(t[-1] for t in ref) >= min_accepts, 'no move was accepted: the run never left its start'
    return worst


def atoms_of(cfg):
    from ase import Atoms
    return Atoms(numbers=cfg['numbers'], positions=cfg['positions'].copy(), cell=cfg['cell'], pbc=cfg['pbc'])


def lj_calc(side, cfg):
    if side == 'ref':
        return OracleLJCalculator(rc=cfg['rc'], sigma=cfg['sigma'], epsilon=cfg['epsilon'])
    from mcpy_b200.calculators import B200Calculator
    return B200Calculator('lj', epsilon=cfg['epsilon'], sigma=cfg['sigma'], rc=cfg['rc'])


def emt_calc(side):
    if side == 'ref':
        return OracleEMTCalculator()
    from mcpy_b200.calculators import B200Calculator
    return B200Calculator('emt')


# ------------------------------------------------------------------------------------------ C1 / C2
def run_box_gcmc(side, cfg, symbol, mu, steps, n_moves=3, min_insert=None):
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    atoms = atoms_of(cfg)
    calc = Logged(lj_calc(side, cfg))
    cell = Cell(atoms, seed=21)
    if side == 'b200' and min_insert is not None:
        from mcpy_b200.moves import B200InsertionMove
        ins = B200InsertionMove(cell, [symbol], 31, calc.inner.engine, min_insert=min_insert)
    else:
        ins = InsertionMove(cell, [symbol], 31, min_insert=min_insert)
    moves = [ins, DeletionMove(cell, [symbol], 32),
             DisplacementMove([symbol], 33, max_displacement=cfg['max_displacement'])]
    ms = MoveSelector([1, 1, 1], moves, seed=34, n_moves=n_moves)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type=cfg['units'], calculator=calc, mu={symbol: mu},
                               species=[symbol], temperature=cfg['temperature'], move_selector=ms, random_seed=35,
                               traj_file=None, outfile=None)
    trace = trace_acceptance(g)
    g.run(steps)
    return trace, g, calc


def test_c1_argon_lj_gcmc():
    """BASELINE config 0: ~500 Ar atoms, metal units, 87 K, ASE LennardJones(eps 0.0104, sigma 3.40, rc 10.2)."""
    cfg = synthetic.s1_argon()
    ref, g_ref, _ = run_box_gcmc('ref', cfg, 'Ar', -0.085, steps=40)
    b200, g_b, calc = run_box_gcmc('b200', cfg, 'Ar', -0.085, steps=40)
    assert_same_traces(ref, b200, min_trials=100, min_accepts=10)
    assert g_ref.n_atoms == g_b.n_atoms and np.array_equal(g_ref.atoms.positions, g_b.atoms.positions)
    assert calc.inner.path_counts[1] > 80          # the run rode the incremental path


@pytest.mark.parametrize('rc', [2.0 ** (1.0 / 6.0), 3.0])
def test_c2_lj_2000_gcmc(rc):
    """BASELINE config 1: 2 000-atom reduced-unit LJ (repulsive WCA cutoff and the attractive rc = 3 variant),
    insert / delete / displace 1:1:1, min_insert through the GPU-backed insertion move."""
    cfg = synthetic.s2_lj(rc=rc)
    ref, g_ref, _ = run_box_gcmc('ref', cfg, 'H', -1.0, steps=30, min_insert=0.8)
    b200, g_b, calc = run_box_gcmc('b200', cfg, 'H', -1.0, steps=30, min_insert=0.8)
    assert_same_traces(ref, b200, min_trials=80, min_accepts=10)
    assert g_ref.n_atoms == g_b.n_atoms and np.array_equal(g_ref.atoms.positions, g_b.atoms.positions)
    assert calc.inner.path_counts[1] > 60


# ------------------------------------------------------------------------------------------ C3
def run_sphere_gcmc(side, steps, n_points=100_000, mu=1.2, T=800.0):
    from ase.cluster import Octahedron
    from mcpy.cell import SphericalCell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    atoms = Octahedron('Ag', 6, 1)
    atoms.set_cell(np.zeros((3, 3)))
    atoms.set_pbc(False)
    calc = Logged(emt_calc(side))
    radii = {'Ag': 2.947, 'O': 0.0}
    if side == 'ref':
        cell = SphericalCell(atoms, vacuum=3.0, species_radii=radii, mc_sample_points=n_points, seed=3)
        ins = InsertionMove(cell, ['O'], 41, min_insert=0.5)
    else:
        from mcpy_b200.cell import B200SphericalCell
        from mcpy_b200.moves import B200InsertionMove
        cell = B200SphericalCell(atoms, vacuum=3.0, species_radii=radii, mc_sample_points=n_points, seed=3,
                                 engine=calc.inner.engine, points='device')
        ins = B200InsertionMove(cell, ['O'], 41, calc.inner.engine, min_insert=0.5)
    moves = [ins, DeletionMove(cell, ['O'], 42), DisplacementMove(['O'], 43, max_displacement=0.5)]
    ms = MoveSelector([2, 1, 1], moves, seed=44, n_moves=4)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='metal', calculator=calc, mu={'O': mu},
                               species=['O'], temperature=T, move_selector=ms, random_seed=45,
                               traj_file=None, outfile=None)
    trace = trace_acceptance(g)
    g.run(steps)
    return trace, g, cell


def test_c3_ag_octahedron_oxygen_emt_spherical_cell():
    """BASELINE config 2: Ag Octahedron(6, 1) + O, ASE EMT, SphericalCell with 100 000 sample points; the free
    volume after every accepted move enters the next insertion / deletion acceptance, so identical traces need
    identical integer counts from the device-regenerated PCG64 stream."""
    ref, g_ref, c_ref = run_sphere_gcmc('ref', steps=30)
    b200, g_b, c_b = run_sphere_gcmc('b200', steps=30)
    assert_same_traces(ref, b200, min_trials=100, min_accepts=10)
    assert g_ref.n_atoms == g_b.n_atoms > 140
    assert c_ref.get_volume() == c_b.get_volume()
    assert c_ref._rng.bit_generator.state == c_b._rng.bit_generator.state     # the two streams stayed in lock-step


# ------------------------------------------------------------------------------------------ C5
def run_molecule_gcmc(side, steps, n_points, mu=1.0, T=800.0):
    from ase import Atoms
    from mcpy.cell import SphericalCell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import (MoleculeDeletionMove, MoleculeDisplacementMove, MoleculeInsertionMove, MoveSelector)
    cfg = synthetic.s5_cupd_co()
    atoms = atoms_of(cfg)
    n_metal = int(np.sum((cfg['numbers'] == 29) | (cfg['numbers'] == 46)))
    ids = np.full(len(atoms), -1, dtype=int)
    ids[n_metal:] = np.repeat(np.arange((len(atoms) - n_metal) // 2), 2)
    atoms.new_array('molecule_id', ids)
    co = Atoms('CO', positions=[[0, 0, 0], [0, 0, 1.128]])
    calc = Logged(emt_calc(side))
    radii = cfg['species_radii']
    if side == 'ref':
        cell = SphericalCell(atoms, vacuum=3.5, species_radii=radii, mc_sample_points=n_points, seed=5)
        moves = [MoleculeInsertionMove(cell, co, 'CO', 51, min_insert=1.3),
                 MoleculeDeletionMove(cell, co, 'CO', 52),
                 MoleculeDisplacementMove(cell, co, 'CO', 53, max_displacement=1.0, max_angle=0.6)]
    else:
        from mcpy_b200.cell import B200SphericalCell
        from mcpy_b200.moves import (B200MoleculeDeletionMove, B200MoleculeDisplacementMove,
                                     B200MoleculeInsertionMove)
        cell = B200SphericalCell(atoms, vacuum=3.5, species_radii=radii, mc_sample_points=n_points, seed=5,
                                 engine=calc.inner.engine, points='device')
        moves = [B200MoleculeInsertionMove(cell, co, 'CO', 51, calc.inner.engine, min_insert=1.3),
                 B200MoleculeDeletionMove(cell, co, 'CO', 52),
                 B200MoleculeDisplacementMove(cell, co, 'CO', 53, max_displacement=1.0, max_angle=0.6)]
    ms = MoveSelector([2, 2, 1], moves, seed=54, n_moves=5)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='metal', calculator=calc, mu={'CO': mu},
                               species=['Cu', 'Pd'], molecules={'CO': co}, temperature=T, move_selector=ms,
                               random_seed=55, traj_file=None, outfile=None)
    trace = trace_acceptance(g)
    g.run(steps)
    return trace, g, cell


def test_c5_cupd_co_molecule_gcmc_emt():
    """BASELINE config 4 (short): 10 425-atom CuPd octahedron + 200 CO, EMT, rigid-molecule insert / delete /
    translate+rotate 2:2:1, SphericalCell free volume (1e6 points in the full configuration; 250 000 here so the
    CPU side of the test stays under a minute)."""
    ref, g_ref, c_ref = run_molecule_gcmc('ref', steps=6, n_points=250_000)
    b200, g_b, c_b = run_molecule_gcmc('b200', steps=6, n_points=250_000)
    assert_same_traces(ref, b200, min_trials=20, min_accepts=3)
    assert g_ref.n_atoms == g_b.n_atoms
    assert np.array_equal(g_ref.atoms.arrays['molecule_id'], g_b.atoms.arrays['molecule_id'])
    assert c_ref.get_volume() == c_b.get_volume()


# ------------------------------------------------------------------------------------------ slab / CustomCell
def run_slab_gcmc(side, steps, mu=1.2, T=800.0):
    from ase.build import fcc111
    from ase.constraints import FixAtoms
    from mcpy.cell import CustomCell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
    atoms = fcc111('Cu', size=(4, 4, 3), vacuum=8.0)
    atoms.set_constraint(FixAtoms(indices=list(range(16))))
    calc = Logged(emt_calc(side))
    top = float(atoms.positions[:, 2].max())
    kw = dict(custom_height=5.0, bottom_z=top + 0.5, species_radii={'Cu': 2.4, 'O': 0.6}, mc_sample_points=50_000, seed=7)
    if side == 'ref':
        cell = CustomCell(atoms, **kw)
    else:
        from mcpy_b200.cell import B200CustomCell
        cell = B200CustomCell(atoms, engine=calc.inner.engine, points='device', **kw)
    moves = [InsertionMove(cell, ['O'], 61), DeletionMove(cell, ['O'], 62),
             DisplacementMove(['O'], 63, max_displacement=0.4)]
    ms = MoveSelector([2, 1, 1], moves, seed=64, n_moves=4)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='metal', calculator=calc, mu={'O': mu},
                               species=['O'], temperature=T, move_selector=ms, random_seed=65,
                               traj_file=None, outfile=None)
    trace = trace_acceptance(g)
    g.run(steps)
    return trace, g, cell


def test_slab_custom_cell_gcmc_emt():
    """A Cu(111) slab (pbc TTF, bottom layer under FixAtoms) + O in a CustomCell: periodic images of the exclusion
    spheres, atoms.wrap() after every accepted move (ULP-level changes of every coordinate under the tracked base),
    constraint rollback on rejected deletions."""
    ref, g_ref, c_ref = run_slab_gcmc('ref', steps=25)
    b200, g_b, c_b = run_slab_gcmc('b200', steps=25)
    assert_same_traces(ref, b200, min_trials=80, min_accepts=8)
    assert g_ref.n_atoms == g_b.n_atoms
    assert c_ref.get_volume() == c_b.get_volume()
    assert [list(c.get_indices()) for c in g_ref.atoms.constraints] == [list(c.get_indices()) for c in g_b.atoms.constraints]


# ------------------------------------------------------------------------------------------ C4
def _c4_factory(calc, n_atoms):
    from mcpy.cell import Cell
    from mcpy.ensembles import GrandCanonicalEnsemble
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector

    def factory(T=None, mu=None, rank=0):
        cfg = synthetic.s2_lj(n=n_atoms, rc=3.0, seed=100 + rank)
        atoms = atoms_of(cfg)
        cell = Cell(atoms, seed=500 + rank)
        moves = [InsertionMove(cell, ['H'], 600 + rank), DeletionMove(cell, ['H'], 700 + rank),
                 DisplacementMove(['H'], 800 + rank, max_displacement=0.3)]
        ms = MoveSelector([1, 1, 1], moves, seed=900 + rank, n_moves=2)
        return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=calc, mu=mu,
                                      species=['H'], temperature=2.0, move_selector=ms, random_seed=1000 + rank,
                                      traj_file=None, outfile=None)
    return factory


def test_c4_batched_replica_exchange_64_x_2000():
    """BASELINE config 3 at its real size: the unmodified BatchedReplicaExchange over 64 replicas of the 2 000-atom
    LJ fluid on a mu ladder, one batched get_potential_energies call per sub-move, exchanges every 2 steps -- with
    the oracle and with B200Calculator (resident bases + one batched delta launch)."""
    from mcpy.ensembles import BatchedReplicaExchange
    mus = [{'H': float(m)} for m in np.linspace(-3.0, -1.8, 64)]
    out = {}
    for side in ('ref', 'b200'):
        calc = Logged(lj_calc(side, synthetic.s2_lj(rc=3.0)))
        re = BatchedReplicaExchange(_c4_factory(calc, 2000), calc, mus=mus, gcmc_steps=5, exchange_interval=2, seed=31,
                                    outfile=None, global_minimum_file=None)
        traces = [trace_acceptance(r) for r in re.replicas]
        re.run()
        out[side] = (re, traces, calc)
    re_r, tr_r, calc_r = out['ref']
    re_b, tr_b, calc_b = out['b200']
    assert re_r.exchange_attempts == re_b.exchange_attempts and re_r.exchange_successes == re_b.exchange_successes
    assert sum(re_r.exchange_successes) > 0
    for a, b in zip(tr_r, tr_b):
        assert_same_traces(a, b, min_trials=1, min_accepts=0)
    assert sum(len(t) for t in tr_r) >= 600
    for a, b in zip(re_r.replicas, re_b.replicas):
        assert a.n_atoms == b.n_atoms and np.array_equal(a.atoms.positions, b.atoms.positions)
        assert abs(a.E_old - b.E_old) <= 1e-10 * max(1.0, abs(a.E_old))
    E_r, E_b = np.array(calc_r.energies), np.array(calc_b.energies)
    assert E_r.shape == E_b.shape and np.all(np.abs(E_r - E_b) <= 1e-10 * np.maximum(1.0, np.abs(E_r)))
    paths = calc_b.inner.path_counts
    # full evaluations: the 64 constructors + the initial re-batch; every trial rode resident bases + one batched delta
    assert paths[0] <= 128 and paths[1] >= 640, paths


# ------------------------------------------------------------------------------------------ canonical / relax
def run_canonical(side, steps):
    from ase.cluster import Octahedron
    from ase.constraints import FixAtoms
    from mcpy.ensembles import CanonicalEnsemble
    from mcpy.moves import DisplacementMove, MoveSelector
    atoms = Octahedron('Cu', 4, 1)
    atoms.set_cell(np.zeros((3, 3)))
    atoms.set_pbc(False)
    rng = np.random.default_rng(9)
    atoms.positions += rng.normal(scale=0.05, size=atoms.positions.shape)
    fixed = FixAtoms(indices=[0, 1, 2])
    atoms.set_constraint(fixed)
    if side == 'ref':
        from ase.calculators.emt import EMT
        from ase.optimize import LBFGS
        calc, opt = EMT(), LBFGS
    else:
        from mcpy_b200.calculators import B200Calculator, B200LBFGS
        calc, opt = B200Calculator('emt', incremental=False), B200LBFGS
    ms = MoveSelector([1], [DisplacementMove(['Cu'], 71, max_displacement=0.6)], seed=72, n_moves=1)
    g = CanonicalEnsemble(atoms=atoms, calculator=calc, random_seed=73, optimizer=opt, fmax=0.05, temperature=800.0,
                          move_selector=ms, constraints=[fixed], traj_file=None, outfile=None)
    energies = []
    inner = g._acceptance_condition

    def logged(diff):
        ok = inner(diff)
        energies.append((float(diff), bool(ok)))
        return ok
    g._acceptance_condition = logged
    g.run(steps)
    return energies, g


def test_canonical_ensemble_relax_loop():
    """a22: the unmodified CanonicalEnsemble (copy -> move -> optimizer(atoms).run(fmax) -> energy, NVT Metropolis)
    with ASE's EMT + LBFGS restated on the CPU, and with B200Calculator + B200LBFGS (the whole relaxation in one
    engine call).  Relaxed energies agree to 1e-7 (two optimiser trajectories in different arithmetic orders, both
    stopped at fmax = 0.05), decisions are identical."""
    ref, g_ref = run_canonical('ref', steps=12)
    b200, g_b = run_canonical('b200', steps=12)
    assert len(ref) == len(b200) >= 10
    for (da, oa), (db, ob) in zip(ref, b200):
        assert oa == ob
        assert abs(da - db) <= 1e-7 * max(1.0, abs(da))
    assert any(o for _, o in ref)
    assert abs(g_ref._current_energy - g_b._current_energy) <= 1e-7 * max(1.0, abs(g_ref._current_energy))
    assert np.allclose(g_ref.atoms.positions, g_b.atoms.positions, atol=1e-5)
    assert np.array_equal(g_ref.atoms.positions[:3], g_b.atoms.positions[:3])      # FixAtoms held exactly
