"""CPU: the LBFGS restatement (oracle/lbfgs.py) on the oracle's own forces -- it reaches a stationary point of the
pinned energies, never ends above the start, holds FixAtoms rows still, and steps=0 is a no-op."""
import numpy as np

from oracle import lbfgs, lj as olj


def _cluster(seed=4, n=13):
    rng = np.random.default_rng(seed)
    g = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], float)[:n] * 1.12
    return g + rng.uniform(-0.08, 0.08, size=g.shape)


def _ef(cell, pbc, rc=2.5):
    def f(p):
        return (olj.lj_energy(p, cell, pbc, 1.0, 1.0, rc), olj.lj_forces(p, cell, pbc, 1.0, 1.0, rc))
    return f


def test_lbfgs_reaches_a_stationary_point_and_lowers_the_energy():
    pos = _cluster()
    cell, pbc = np.diag([20.0] * 3), np.zeros(3, bool)
    ef = _ef(cell, pbc)
    e0 = ef(pos)[0]
    e, p, steps, fm = lbfgs.lbfgs_relax(ef, pos, fmax=1e-3, steps=500)
    assert 0 < steps < 500 and fm < 1e-3 and e < e0
    assert np.linalg.norm(olj.lj_forces(p, cell, pbc, 1.0, 1.0, 2.5), axis=1).max() < 1e-3
    # no atom moves farther than maxstep in one step: the first step from rest is exactly maxstep long at most
    e1, p1, s1, _ = lbfgs.lbfgs_relax(ef, pos, fmax=1e-3, steps=1)
    assert s1 == 1 and np.sqrt(((p1 - pos) ** 2).sum(1)).max() <= 0.2 + 1e-12


def test_steps_zero_and_fixed_rows():
    pos = _cluster(seed=5)
    cell, pbc = np.diag([20.0] * 3), np.zeros(3, bool)
    ef = _ef(cell, pbc)
    e, p, steps, _ = lbfgs.lbfgs_relax(ef, pos, fmax=1e-3, steps=0)
    assert steps == 0 and np.array_equal(p, pos) and e == ef(pos)[0]
    e, p, steps, fm = lbfgs.lbfgs_relax(ef, pos, fmax=1e-3, steps=300, fixed=[0, 3, 7])
    assert np.array_equal(p[[0, 3, 7]], pos[[0, 3, 7]]) and fm < 1e-3 and not np.array_equal(p[1], pos[1])
