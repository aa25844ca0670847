"""GPU: min_insert distance test (K9) and the insertion moves built on it.

With the reference checkout present the moves are also run side by side with the unmodified
InsertionMove / MoleculeInsertionMove: same seeds -> same inserted positions, same RNG streams."""
import os
import sys

import numpy as np
import pytest

from mcpy_b200 import Engine

pytestmark = pytest.mark.gpu


def _mic_min_d2(points, pos, cell_len, mask):
    out = np.full(len(points), np.inf)
    for c, p in enumerate(points):
        d = pos[mask] - p
        d -= cell_len * np.round(d / cell_len)
        out[c] = np.einsum('ij,ij->i', d, d).min()
    return out


def test_min_dist2_matches_minimum_image(lj_engine):
    rng = np.random.default_rng(2)
    L = np.array([12.0, 13.0, 11.0])
    pos = rng.random((300, 3)) * L
    mask = rng.random(300) < 0.6
    lj_engine.set_lj(1.0, 1.0, 3.0)
    lj_engine.upload([pos], [np.ones(300, np.int32)], np.diag(L).reshape(1, 9), np.ones((1, 3), bool))
    pts = rng.random((500, 3)) * L
    got = lj_engine.min_dist2(pts, 0, mask)
    ref = _mic_min_d2(pts, pos, L, mask)
    near = ref <= 9.0
    np.testing.assert_allclose(got[near], ref[near], rtol=1e-13)
    assert np.all(np.isinf(got[~near]))
    assert np.array_equal(lj_engine.min_dist2(pts, 0, None) <= got, np.ones(500, bool))


def _words(rng):
    st = rng.bit_generator.state['state']
    m = (1 << 64) - 1
    return [st['state'] >> 64, st['state'] & m, st['inc'] >> 64, st['inc'] & m]


def test_insertion_moves_follow_the_reference_stream(lj_engine):
    """tests/golden/insertion_trace.npz was recorded from the UNMODIFIED InsertionMove /
    MoleculeInsertionMove; same seeds -> same inserted positions and the same generator states."""
    from ase import Atoms
    from mcpy.cell import Cell
    from mcpy_b200.moves import B200InsertionMove, B200MoleculeInsertionMove

    def BoxCell(atoms, species_radii, seed):       # the reference's whole-box cell
        return Cell(atoms, species_radii=species_radii, seed=seed)

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'insertion_trace.npz'))
    lj_engine.set_lj(1.0, 1.0, 3.0)
    L = float(d['L'])
    a = Atoms('H150', positions=d['base'], cell=[L, L, L], pbc=True)
    c = BoxCell(a, {'H': 1.0}, seed=9)
    m = B200InsertionMove(c, species=['H'], seed=5, engine=lj_engine, min_insert=1.15)
    m.batch = 7                                    # force several batches per proposal
    for k in range(12):
        r = m.do_trial_move(a)
        assert (r[0] is not False) == bool(d['atom_ok'][k])
        assert np.array_equal(a.positions, d['atom_pos'][d['atom_off'][k]:d['atom_off'][k + 1]])
        assert _words(c._rng) == [int(x) for x in d['atom_cell_state'][k]]
    dimer = Atoms('H2', positions=[[0, 0, 0], [0, 0, 0.9]])
    a = Atoms('H150', positions=d['base'], cell=[L, L, L], pbc=True)
    c = BoxCell(a, {'H': 1.0}, seed=19)
    m = B200MoleculeInsertionMove(c, dimer, 'H2', seed=6, engine=lj_engine, min_insert=1.1)
    m.batch = 5
    for k in range(8):
        r = m.do_trial_move(a)
        assert (r[0] is not False) == bool(d['mol_ok'][k])
        assert np.array_equal(a.positions, d['mol_pos'][d['mol_off'][k]:d['mol_off'][k + 1]])
        assert _words(c._rng) == [int(x) for x in d['mol_cell_state'][k]]
        assert np.array_equal(np.array(m.rng.random.getstate()[1], dtype=np.uint64), d['mol_py_state'][k])
        assert (-1 if m.last_exchange_count is None else m.last_exchange_count) == int(d['mol_count'][k])
    assert np.array_equal(np.asarray(a.arrays['molecule_id']), d['mol_ids'])
