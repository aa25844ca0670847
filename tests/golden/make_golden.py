"""Generate golden fixtures by running the UNMODIFIED reference drivers (farrisric/mcpy at
/root/reference) in the build container, with the CPU oracle as calculator.

    python tests/golden/make_golden.py

The reference needs ``ase`` (absent here, no network); the test-only stand-in under
``baseline/ase_shim`` supplies ``Atoms`` etc.  Everything numeric that ends up in the fixtures
comes from the reference's own code paths (moves, cells, ``GrandCanonicalEnsemble``,
numpy PCG64 / Python MT19937 streams, scipy cKDTree) and from the oracle energies.

Fixtures (small, committed):
  lj_gcmc_trace.npz   -- 240 trials of LJ GCMC in a periodic box (reference Cell + Insertion /
                         Deletion / Displacement moves, LJ units): every configuration handed to
                         compute_energy, the oracle energy, the acceptance inputs, the uniform
                         draws and the accept / reject decision.
  emt_sphere_trace.npz -- 220 trials of Ag octahedron + O GCMC with EMT in a SphericalCell:
                         the same, plus every free-volume refresh (PCG64 state before, atoms,
                         resulting volume).
  insertion_trace.npz -- reference InsertionMove / MoleculeInsertionMove with min_insert.
  molecule_moves_trace.npz -- reference MoleculeDisplacementMove (free and capped rotation) and
                         MoleculeDeletionMove on dimers in a periodic box and in a SphericalCell,
                         with atoms.wrap() between moves: configuration, molecule ids, Python
                         MT19937 state, outcome and last_exchange_count after every proposal.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baseline import refpath  # noqa: E402
ASE_KIND, MCPY_ROOT = refpath.enable()

from ase import Atoms  # noqa: E402
from ase.cluster import Octahedron  # noqa: E402
from mcpy.cell import Cell, SphericalCell  # noqa: E402
from mcpy.ensembles import GrandCanonicalEnsemble  # noqa: E402
from mcpy.moves import (DeletionMove, DisplacementMove, InsertionMove, MoleculeInsertionMove,  # noqa: E402
                        MoleculeDeletionMove, MoleculeDisplacementMove, MoveSelector)
from oracle import cfast  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class Recorder:
    """Calculator that records every configuration it is asked about."""

    def __init__(self, fn):
        self.fn = fn
        self.pos, self.num, self.energy = [], [], []

    def get_potential_energy(self, atoms):
        e = self.fn(atoms)
        self.pos.append(atoms.positions.copy())
        self.num.append(atoms.numbers.copy())
        self.energy.append(e)
        return e


def instrument(g):
    """Record the acceptance inputs, draws and outcome of every trial."""
    log = {'dE': [], 'dN': [], 'volume': [], 'species': [], 'n': [], 'accept': [], 'draw': []}
    orig_cond = g._acceptance_condition
    orig_uniform = g.rng_acceptance.get_uniform
    state = {'draw': np.nan}

    def uniform(a=0.0, b=1.0):
        u = orig_uniform(a, b)
        state['draw'] = u
        return u

    def cond(dE, dN, volume, species, n_atoms=None):
        state['draw'] = np.nan
        ok = orig_cond(dE, dN, volume, species, n_atoms)
        for k, v in zip(('dE', 'dN', 'volume', 'species', 'n', 'accept', 'draw'),
                        (dE, dN, volume, species, n_atoms, bool(ok), state['draw'])):
            log[k].append(v)
        return ok

    g.rng_acceptance.get_uniform = uniform
    g._acceptance_condition = cond
    return log


def pack(rec):
    off = np.concatenate([[0], np.cumsum([len(p) for p in rec.pos])]).astype(np.int64)
    return dict(cfg_off=off, cfg_pos=np.concatenate(rec.pos), cfg_num=np.concatenate(rec.num).astype(np.int32),
                cfg_energy=np.array(rec.energy))


def lj_trace():
    L, rc, T, mu = 9.0, 3.0, 2.0, -2.4       # the reference's LAMMPS-parity box (benchmark/lammps_gcmc_parity.py)
    rng = np.random.default_rng(5)
    atoms = Atoms('H60', positions=rng.uniform(0.9, L - 0.9, size=(60, 3)), cell=[L, L, L], pbc=True)
    rec = Recorder(lambda a: cfast.lj_energy(a.positions, np.asarray(a.cell), a.pbc, 1.0, 1.0, rc))
    cell = Cell(atoms, seed=7)
    moves = [InsertionMove(cell, species=['H'], seed=21), DeletionMove(cell, species=['H'], seed=22, min_atoms=1),
             DisplacementMove(species=['H'], seed=23, max_displacement=0.3)]
    ms = MoveSelector([1, 1, 1], moves, seed=14, n_moves=1)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=rec, mu={'H': mu},
                               species=['H'], temperature=T, move_selector=ms, random_seed=13,
                               traj_file=None, outfile=None)
    log = instrument(g)
    for _ in range(240):
        g.do_gcmc_step()
    d = pack(rec)
    d.update(cell=np.asarray(atoms.cell), pbc=np.asarray(atoms.pbc), rc=rc, temperature=T, mu=mu,
             dE=np.array(log['dE']), dN=np.array(log['dN']), volume=np.array(log['volume']),
             n=np.array(log['n']), accept=np.array(log['accept']), draw=np.array(log['draw']),
             final_energy=g.E_old, final_n=len(g.atoms))
    np.savez_compressed(os.path.join(OUT, 'lj_gcmc_trace.npz'), **d)
    print('lj trace:', len(rec.energy), 'energy calls,', int(np.sum(log['accept'])), 'accepted, final N', len(g.atoms))


def emt_trace():
    atoms = Octahedron('Ag', 4, 1)            # the reference's equivalence-test system
    radii = {'Ag': 2.947, 'O': 0.0}
    scell = SphericalCell(atoms, vacuum=3.0, species_radii=radii, mc_sample_points=20_000, seed=17)
    vol_log = {'state': [], 'pos': [], 'num': [], 'volume': []}
    orig_calc = scell.calculate_volume

    def calc_volume(a):
        st = scell._rng.bit_generator.state['state']
        m = (1 << 64) - 1
        vol_log['state'].append([st['state'] >> 64, st['state'] & m, st['inc'] >> 64, st['inc'] & m])
        vol_log['pos'].append(a.positions.copy())
        vol_log['num'].append(a.numbers.copy())
        orig_calc(a)
        vol_log['volume'].append(scell.volume)

    scell.calculate_volume = calc_volume
    rec = Recorder(lambda a: cfast.emt_energy(a.positions, a.numbers, np.asarray(a.cell), a.pbc))
    moves = [InsertionMove(scell, species=['O'], seed=31, min_insert=0.5),
             DeletionMove(scell, species=['O'], seed=32),
             DisplacementMove(species=['O'], seed=33, max_displacement=0.25,
                              constraints=list(range(len(atoms))))]
    ms = MoveSelector([2, 1, 1], moves, seed=34, n_moves=1)
    g = GrandCanonicalEnsemble(atoms=atoms, cells=[scell], units_type='metal', calculator=rec, mu={'O': 1.2},
                               species=['O'], temperature=800.0, move_selector=ms, random_seed=35,
                               traj_file=None, outfile=None)
    log = instrument(g)
    for _ in range(220):
        try:
            g.do_gcmc_step()
        except ValueError:
            # DisplacementMove with no movable O yet raises in the reference; skip such steps
            continue
    d = pack(rec)
    voff = np.concatenate([[0], np.cumsum([len(p) for p in vol_log['pos']])]).astype(np.int64)
    d.update(cell=np.asarray(atoms.cell), pbc=np.asarray(atoms.pbc), temperature=800.0, mu=1.2,
             dE=np.array(log['dE']), dN=np.array(log['dN']), volume=np.array(log['volume']),
             n=np.array(log['n']), accept=np.array(log['accept']), draw=np.array(log['draw']),
             lambda_O=g.units.lambda_dbs['O'], beta=g.units.beta,
             sphere_radius=scell.radius, n_points=scell.mc_sample_points,
             vol_off=voff, vol_pos=np.concatenate(vol_log['pos']),
             vol_num=np.concatenate(vol_log['num']).astype(np.int32),
             vol_state=np.array(vol_log['state'], dtype=np.uint64), vol_value=np.array(vol_log['volume']),
             final_energy=g.E_old, final_n=len(g.atoms))
    np.savez_compressed(os.path.join(OUT, 'emt_sphere_trace.npz'), **d)
    print('emt trace:', len(rec.energy), 'energy calls,', int(np.sum(log['accept'])), 'accepted,',
          len(vol_log['volume']), 'volume refreshes, final N', len(g.atoms))


def _pcg_words(rng):
    st = rng.bit_generator.state['state']
    m = (1 << 64) - 1
    return [st['state'] >> 64, st['state'] & m, st['inc'] >> 64, st['inc'] & m]


def insertion_trace():
    """Reference InsertionMove / MoleculeInsertionMove with min_insert in a dense periodic box:
    configuration and generator states after every proposal."""
    L = 8.0
    rng = np.random.default_rng(4)
    base = Atoms('H150', positions=rng.uniform(0, L, size=(150, 3)), cell=[L, L, L], pbc=True)
    out = dict(L=L, base=base.positions.copy())
    a = base.copy()
    c = Cell(a, species_radii={'H': 1.0}, seed=9)
    m = InsertionMove(c, species=['H'], seed=5, min_insert=1.15)
    pos, ok, states = [], [], []
    for _ in range(12):
        r = m.do_trial_move(a)
        ok.append(r[0] is not False)
        pos.append(a.positions.copy())
        states.append(_pcg_words(c._rng))
    out.update(atom_off=np.concatenate([[0], np.cumsum([len(p) for p in pos])]), atom_pos=np.concatenate(pos),
               atom_ok=np.array(ok), atom_cell_state=np.array(states, dtype=np.uint64))
    dimer = Atoms('H2', positions=[[0, 0, 0], [0, 0, 0.9]])
    a = base.copy()
    c = Cell(a, species_radii={'H': 1.0}, seed=19)
    m = MoleculeInsertionMove(c, dimer, 'H2', seed=6, min_insert=1.1)
    pos, ok, states, pystates, counts = [], [], [], [], []
    for _ in range(8):
        r = m.do_trial_move(a)
        ok.append(r[0] is not False)
        pos.append(a.positions.copy())
        states.append(_pcg_words(c._rng))
        st = m.rng.random.getstate()
        pystates.append(np.array(st[1], dtype=np.uint64))
        counts.append(-1 if m.last_exchange_count is None else m.last_exchange_count)
    out.update(mol_off=np.concatenate([[0], np.cumsum([len(p) for p in pos])]), mol_pos=np.concatenate(pos),
               mol_ok=np.array(ok), mol_cell_state=np.array(states, dtype=np.uint64),
               mol_py_state=np.array(pystates), mol_count=np.array(counts),
               mol_ids=np.asarray(a.arrays['molecule_id']))
    np.savez_compressed(os.path.join(OUT, 'insertion_trace.npz'), **out)
    print('insertion trace:', int(np.sum(out['atom_ok'])), 'of 12 atoms,', int(np.sum(out['mol_ok'])), 'of 8 dimers placed')


def molecule_system(kind):
    """60 atoms + 24 dimers; 'box': periodic cube with the whole-box Cell, 'sphere': open cluster in a
    SphericalCell.  Returned together with the cell (shared by the generator and the replay test)."""
    rng = np.random.default_rng(12)
    L = 9.0
    free = rng.uniform(0, L, size=(60, 3))
    centres = rng.uniform(0, L, size=(24, 3))
    axes = rng.normal(size=(24, 3))
    axes /= np.linalg.norm(axes, axis=1)[:, None]
    pos = [free]
    ids = [np.full(60, -1)]
    for m in range(24):
        pos.append(np.array([centres[m] - 0.45 * axes[m], centres[m] + 0.45 * axes[m]]))
        ids.append(np.full(2, m))
    pos = np.concatenate(pos)
    symbols = ['H'] * 60 + ['C', 'O'] * 24
    if kind == 'box':
        a = Atoms(symbols, positions=pos, cell=[L, L, L], pbc=True)
        a.wrap()
        a.new_array('molecule_id', np.concatenate(ids).astype(int))
        c = Cell(a, species_radii={'H': 1.0, 'C': 0.0, 'O': 0.0}, seed=21)
    else:
        a = Atoms(symbols, positions=pos - pos.mean(axis=0), cell=[0, 0, 0], pbc=False)
        a.new_array('molecule_id', np.concatenate(ids).astype(int))
        c = SphericalCell(a, vacuum=-3.0, species_radii={'H': 1.0, 'C': 0.0, 'O': 0.0}, mc_sample_points=1000, seed=22)
    return a, c


def molecule_moves_trace():
    out = {}
    co = Atoms('CO', positions=[[0, 0, 0], [0, 0, 0.9]])
    for kind in ('box', 'sphere'):
        a, c = molecule_system(kind)
        moves = [MoleculeDisplacementMove(c, co, 'CO', seed=31, max_displacement=0.6),
                 MoleculeDisplacementMove(c, co, 'CO', seed=32, max_displacement=0.4, max_angle=0.6),
                 MoleculeDeletionMove(c, co, 'CO', seed=33, min_molecules=3)]
        order = np.random.default_rng(7).integers(0, 3, size=60)
        order[order == 2] = np.where(np.random.default_rng(8).random(np.sum(order == 2)) < 0.5, 2, 0)
        pos, off, ok, states, counts, idsout = [], [0], [], [], [], []
        for k in order:
            m = moves[int(k)]
            r = m.do_trial_move(a)
            ok.append(r[0] is not False)
            if r[0] is not False and kind == 'box':
                a.wrap()                                   # accepted periodic moves are wrapped by the ensemble
            pos.append(a.positions.copy())
            off.append(off[-1] + len(a))
            idsout.append(np.asarray(a.arrays['molecule_id']).copy())
            states.append(np.array(m.rng.random.getstate()[1], dtype=np.uint64))
            lc = getattr(m, 'last_exchange_count', None)
            counts.append(-1 if lc is None else lc)
        out.update({f'{kind}_order': order, f'{kind}_pos': np.concatenate(pos), f'{kind}_off': np.array(off),
                    f'{kind}_ok': np.array(ok), f'{kind}_state': np.array(states),
                    f'{kind}_count': np.array(counts), f'{kind}_ids': np.concatenate(idsout)})
        print(f'molecule trace ({kind}):', int(np.sum(ok)), 'of', len(order), 'proposals made;',
              len(a), 'atoms left')
    np.savez_compressed(os.path.join(OUT, 'molecule_moves_trace.npz'), **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['lj', 'emt', 'insertion', 'molecules']
    if 'lj' in which:
        lj_trace()
    if 'emt' in which:
        emt_trace()
    if 'insertion' in which:
        insertion_trace()
    if 'molecules' in which:
        molecule_moves_trace()
