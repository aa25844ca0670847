"""CPU: the rigid-molecule deletion / displacement mirrors replay a trace recorded from the UNMODIFIED
reference classes (tests/golden/make_golden.py::molecule_moves_trace): same seeds -> the same proposals,
positions, molecule ids, MT19937 states and last_exchange_count, bit for bit -- while the candidate
search re-evaluates only the molecules that moved."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _system(kind):
    """tests/golden/make_golden.py::molecule_system, with duck-typed cells instead of the reference's."""
    from ase import Atoms
    rng = np.random.default_rng(12)
    L = 9.0
    free = rng.uniform(0, L, size=(60, 3))
    centres = rng.uniform(0, L, size=(24, 3))
    axes = rng.normal(size=(24, 3))
    axes /= np.linalg.norm(axes, axis=1)[:, None]
    pos, ids = [free], [np.full(60, -1)]
    for m in range(24):
        pos.append(np.array([centres[m] - 0.45 * axes[m], centres[m] + 0.45 * axes[m]]))
        ids.append(np.full(2, m))
    pos = np.concatenate(pos)
    symbols = ['H'] * 60 + ['C', 'O'] * 24

    class Box:                                      # mcpy.cell.Cell: the whole periodic box
        def is_point_inside(self, p):
            return True
        is_point_exchangeable = is_point_inside

    class Sphere:                                   # SphericalCell.is_point_inside (spherical_cell.py:48-53)
        def __init__(self, center, radius):
            self.center, self.radius = center, radius

        def is_point_inside(self, p):
            diff = p - self.center
            return float(np.dot(diff, diff)) <= self.radius ** 2
        is_point_exchangeable = is_point_inside

    if kind == 'box':
        a = Atoms(symbols, positions=pos, cell=[L, L, L], pbc=True)
        a.wrap()
        a.new_array('molecule_id', np.concatenate(ids).astype(int))
        return a, Box()
    a = Atoms(symbols, positions=pos - pos.mean(axis=0), cell=[0, 0, 0], pbc=False)
    a.new_array('molecule_id', np.concatenate(ids).astype(int))
    # SphericalCell.__init__ (spherical_cell.py:30-46): the atoms are translated so that their centre of
    # mass sits at the origin; radius = largest distance from it + vacuum
    center = np.zeros(3)
    a.translate(center - a.get_center_of_mass())
    radius = float(np.linalg.norm(a.positions - center, axis=1).max() + (-3.0))
    return a, Sphere(center, radius)


@pytest.mark.parametrize('kind', ['box', 'sphere'])
def test_molecule_moves_follow_the_reference_trace(kind):
    from ase import Atoms
    from mcpy_b200.moves import B200MoleculeDeletionMove, B200MoleculeDisplacementMove
    d = np.load(os.path.join(HERE, 'golden', 'molecule_moves_trace.npz'))
    a, c = _system(kind)
    co = Atoms('CO', positions=[[0, 0, 0], [0, 0, 0.9]])
    moves = [B200MoleculeDisplacementMove(c, co, 'CO', seed=31, max_displacement=0.6),
             B200MoleculeDisplacementMove(c, co, 'CO', seed=32, max_displacement=0.4, max_angle=0.6),
             B200MoleculeDeletionMove(c, co, 'CO', seed=33, min_molecules=3)]
    off, ioff = d[f'{kind}_off'], 0
    for k, which in enumerate(d[f'{kind}_order']):
        m = moves[int(which)]
        r = m.do_trial_move(a)
        assert (r[0] is not False) == bool(d[f'{kind}_ok'][k]), k
        if r[0] is not False and kind == 'box':
            a.wrap()
        assert np.array_equal(a.positions, d[f'{kind}_pos'][off[k]:off[k + 1]]), k
        assert np.array_equal(np.asarray(a.arrays['molecule_id']), d[f'{kind}_ids'][off[k]:off[k + 1]])
        assert np.array_equal(np.array(m.rng.random.getstate()[1], dtype=np.uint64), d[f'{kind}_state'][k])
        lc = getattr(m, 'last_exchange_count', None)
        assert (-1 if lc is None else lc) == int(d[f'{kind}_count'][k])
    # the index re-evaluated only what moved: far fewer predicate evaluations than
    # (molecules x proposals), which is what the reference's find_molecules does
    n_eval = sum(m.index.recomputed for m in moves)
    assert n_eval < 24 * 60 / 3


def test_index_matches_find_molecules_semantics():
    """Composition filter, ascending molecule-id order, ascending member order, -1 ignored."""
    from ase import Atoms
    from mcpy_b200.moves.molecules import MoleculeIndex
    a = Atoms(['O', 'H', 'C', 'C', 'O', 'H', 'O', 'C'],
              positions=np.arange(24, dtype=float).reshape(8, 3), cell=[30, 30, 30], pbc=False)
    a.new_array('molecule_id', np.array([5, -1, 5, 2, 2, 7, 9, -1]))
    idx = MoleculeIndex(sorted(['C', 'O']))
    got = idx.groups(a, None)
    assert [list(g) for g in got] == [[3, 4], [0, 2]]            # id 2 before id 5; H-only id 7 and lone O id 9 dropped

    class Half:
        def is_point_inside(self, p):
            return p[0] < 8.0
    got = idx.groups(a, Half())
    assert [list(g) for g in got] == [[0, 2]]                    # COM of (3,4) has x > 8
